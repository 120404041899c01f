#!/usr/bin/env python
"""bench.py -- sampled SMILES/s (BASELINE cfg 2) and train tokens/s (cfg 3 / cfg 4) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic batch:
  * sampling (headline, cfg 2): one sample_smiles call = one batch of --batch latent draws (default 30000 = cfg 2's 30k draws
    in one call; the reference driver's -batch_size default 512 is reported under "batch512") decoded to max_strlen=100
    (99 KV-cached multinomial steps), vaetf.  `value`: inputs resident in HBM, CUDA events.  `e2e`: the public call with
    HOST buffers (pinned latents + lengths in, Python strings out).  `e2e_public_call`: sample_smiles(n) with NO inputs
    (lengths and latents drawn inside the call) for the reference-faithful host draw and for z_on_device=True;
  * "cfg5": scavaetf scaffold-conditioned sampling (20-token scaffold prefix), --cfg5-draws per rank;
  * "train": one optimiser step (fwd + loss + bwd [+ NCCL all-reduce of the gradients] + Adam) of cfg 3 (pvaetf B=512 S=78, N=1
    only) and cfg 4 (pscavaetf B=512/GPU S=98 data-parallel, every N including 1; also the authors' batch 64), each with its
    own tensor roofline; at N>1 "dp_parity" compares the N-rank gradients with a 1-rank run of the same global batch;
  * "gpu_eager_baseline": the oracle port (= the reference's arithmetic) run eagerly in fp32 on the same GPU, TF32 off / on.
`--impl reference` times the CPU oracle port of the reference's un-cached sampler on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ARCH = dict(N=6, d_model=512, dff=2048, h=8, latent_dim=128)
VOCAB = 32
MAX_STRLEN = 100
BATCH = 512
# dram__bytes_read.sum + dram__bytes_write.sum of one damma::decode_attn_mma_kernel<2> launch (ncu --set full, B=30000 = the rows
# per launch of the roofline leg and of the benched decode, 49 cached keys, profiles/r02_decode_attn_b30000_t49_v7_ncu_raw.csv:
# 3106.2 MB read + 98.0 MB written in 489.2 us) next to the algorithmic bytes of that same launch
NCU_TRAFFIC = {"dram_bytes_per_launch": 3204180864, "algorithmic_bytes_same_launch": 3194880000, "shape": "B=30000, 49 cached keys + 1 new",
               "source": "profiles/r02_decode_attn_b30000_t49_v7_ncu_details.txt"}
SCAFFOLD20 = "c1ccc(cc1)C(=O)NCCOC"        # 20 tokens under the character tokeniser below (cfg 5)
ITOS = ["<unk>", "<pad>", "<sos>", "<eos>", "<sep>"] + list("CcNnOoSsFIBrl()[]=#123456+-H@/")[:27]


class _Vocab:
    def __init__(self):
        self.itos = ITOS
        self.stoi = {t: i for i, t in enumerate(ITOS)}

    def __len__(self):
        return len(self.itos)


class _Field:
    batch_first = True

    def __init__(self):
        self.vocab = _Vocab()

    def tokenize(self, smi):
        return list(smi)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


def toklen_data():
    """Stand-in for the absent toklen_list.csv: round(N(35,7^2)) clipped to [13,55] (BASELINE.md cfg 2)."""
    rng = np.random.RandomState(7)
    return np.clip(np.rint(rng.normal(35, 7, size=20000)), 13, 55)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.p, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================================
# our arm
# =============================================================================================
def build_sampler(dev, dtype="bf16", model_type="vaetf", latent_bucket=8):          # 8 = the sampler's own default
    from gct_plus_b200.Model import Cvaetf, Vaetf
    from gct_plus_b200.Inference.sampling_tool import sampling_tool_dict
    torch.manual_seed(0)
    cls = Vaetf if model_type == "vaetf" else Cvaetf
    model = cls(VOCAB, VOCAB, dropout=0.1, nconds=0, compute_dtype=dtype, **ARCH).to(dev).eval()
    kwargs = dict(top_k=None, latent_dim=ARCH["latent_dim"], max_strlen=MAX_STRLEN, use_cond2dec=False, decode_algo="multinomial",
                  n_jobs=1, toklen_data=toklen_data(), cond_dim=0, scaler=None, device=dev, SRC=_Field(), TRG=_Field(),
                  sync_every=33, latent_bucket=latent_bucket)
    return sampling_tool_dict[model_type](model, kwargs)


def sample_inputs(sampler, n_batches, seed, pinned=True):
    """Per-batch latent lengths and N(0,1) latents on the HOST (pinned), like the reference's sample_z."""
    np.random.seed(seed)
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        toklen = sampler.sample_toklen(BATCH)
        Lz = int(max(toklen))
        zs = torch.randn(BATCH, Lz, ARCH["latent_dim"], generator=g)
        out.append((toklen, zs.pin_memory() if pinned else zs))
    return out


def decode_alg_bytes(cfg_layers, B, Sm, steps, esize, d=512, dff=2048, latent_form=False, lat=128, heads=8):
    """sum over steps of KV read + KV write + weight read.  latent_form: the cross-attention term of this
    build's algorithm (decode_zattn.cuh): per layer the latent rows [Sm, lat] + the folded query / context vectors
    [heads*lat] instead of the memory's K and V projections [Sm, d] x 2; the folded q / out weights are [heads*lat, d]."""
    cross = (Sm * lat + 2 * heads * lat) if latent_form else 2 * Sm * d
    kv = sum(cfg_layers * B * (2 * (t + 1) * d + cross) * esize for t in range(steps))
    kvw = steps * cfg_layers * B * 2 * d * esize
    wq = (6 * d * d + 2 * heads * lat * d) if latent_form else 8 * d * d
    wr = steps * cfg_layers * (wq + 2 * d * dff) * esize
    return kv + kvw + wr


def time_decode_attention(sampler, dev, steps, Sm):
    """Roofline leg: the KV-cache attention kernel alone, one launch per (step, layer), self-attention shapes of a real
    decode (n_cached = t), caches of all layers in rotation (6 x 2 x 52 MB bf16 > L2).  CUDA events on the launch stream."""
    import gct_plus_b200._lib as L
    lib = L.lib()
    d, H, N, B = 512, 8, 6, BATCH                   # rows per launch == the benched batch; 30000 x 6 layers x 2 x 100 keys = 37 GB of cache
    Lmax = steps + 1
    kc = torch.randn(N, B, Lmax, d, device=dev).bfloat16()
    vc = torch.randn(N, B, Lmax, d, device=dev).bfloat16()
    qkv = torch.randn(B, 3 * d, device=dev).bfloat16()
    out = torch.empty(B, d, device=dev, dtype=torch.bfloat16)
    valid = torch.ones(B, Lmax, device=dev, dtype=torch.uint8)
    def launch(l, t):
        L.check(lib.gct_decode_attention(L.ptr(qkv), 3 * d, qkv[:, d:].data_ptr(), qkv[:, 2 * d:].data_ptr(), 3 * d,
                                         kc[l].data_ptr(), vc[l].data_ptr(), Lmax * d, d, t, L.ptr(valid), Lmax, L.ptr(out), d, B, H,
                                         L.DTYPE_BF16, L.stream_ptr()))
    for t in (10, 50, 90):
        for l in range(N):
            launch(l, t)
    torch.cuda.synchronize()
    # replayed from a CUDA graph so that the python/ctypes launch cost does not pace the GPU
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in range(steps):
            for l in range(N):
                launch(l, t)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = steps * N
    # algorithmic bytes of one launch at step t: read t cached K and V rows + this step's q,k,v, write k,v rows and the output
    alg = sum(B * (2 * t * d + 3 * d + 2 * d + d) * 2 for t in range(steps)) * N
    return alg / launches, ms / launches, launches, B


def run_sampling(args, rank, world, dev):
    import gct_plus_b200._lib as L
    sampler = build_sampler(dev)
    steps = MAX_STRLEN - 1
    K, W = args.steps, args.warmup
    inputs = sample_inputs(sampler, K + W, seed=1000 + rank)
    # ---- device-resident leg (value) ----
    dev_in = []
    for toklen, zs in inputs:
        Lz = zs.size(1)
        mask = (torch.arange(Lz).expand(BATCH, 1, Lz) < torch.LongTensor(toklen).view(BATCH, 1, 1))
        dev_in.append((zs.to(dev), mask.to(dev)))
    ys0 = torch.full((BATCH, 1), 2, dtype=torch.long, device=dev)
    torch.cuda.synchronize()
    for i in range(W):
        sampler.decode(zs=dev_in[i][0], ys=ys0, src_mask=dev_in[i][1])
    clocks = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else rank)
    barrier(world)
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_steps_run = 0
    for i in range(W, W + K):
        sampler.decode(zs=dev_in[i][0], ys=ys0, src_mask=dev_in[i][1])
        n_steps_run += sampler.last_decode_steps
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ck = clocks.stop()
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    value = world * K * BATCH / (ms / 1e3)
    # ---- end-to-end leg through the public API: pinned host latents in, strings out ----
    for i in range(min(W, 2)):
        sampler.sample_smiles(BATCH, zs=inputs[i][1], toklen=list(inputs[i][0]))
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_out = 0
    for i in range(W, W + K):
        smiles, _, _ = sampler.sample_smiles(BATCH, zs=inputs[i][1], toklen=list(inputs[i][0]))
        n_out += len(smiles)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    assert n_out == K * BATCH
    # ---- the public call with NO inputs: sample_smiles(n) draws the lengths (host, NumPy stream) and the latents itself.
    # z on the host = the reference's torch.normal on the CPU generator (seed-faithful, sampling_tool.py:93-97);
    # z_on_device=True (sampler kwarg) draws them with the CUDA generator.
    public = {}
    if getattr(args, "public_call", True):
        for name, on_dev, reps in (("z_on_device", True, max(2, min(K, 4))), ("z_on_host", False, 1)):
            sampler.z_on_device = on_dev
            np.random.seed(4242 + rank)
            torch.manual_seed(4242 + rank)
            sampler.sample_smiles(BATCH)                                  # warm-up (static buffers of this shape exist already)
            barrier(world)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = 0
            for _ in range(reps):
                got += len(sampler.sample_smiles(BATCH)[0])
            torch.cuda.synchronize()
            dt = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
            assert got == reps * BATCH
            public[name] = {"value": world * got / (dt / 1e3), "unit": "SMILES/s", "calls": reps, "ms_per_call": dt / reps}
        sampler.z_on_device = False
        public["note"] = ("sampler.sample_smiles(n) exactly as the reference drivers call it (uc_sampling.py:16-23): sample_toklen + sample_z "
                          "inside the timed call; z_on_host draws 30000 x Lz x 128 normals with torch's CPU generator like the reference")
    # ---- active-row decode (sampler kwarg skip_finished=True; NOT the headline: `value` / `e2e` above decode every row at every
    # step like the reference's loop).  Same inputs; rows that have emitted <eos> stop costing attention work and are gathered out of
    # the batch, so the cost follows the sum of the generated lengths instead of rows x steps.  How much that saves depends on
    # when rows emit <eos>: with these random-init weights ~1 / vocabulary per step.
    active = None
    if getattr(args, "active_rows", True) and BATCH >= 4096:
        sampler.skip_finished = True
        reps = max(2, min(K, 4))
        for i in range(2):
            sampler.decode(zs=dev_in[i][0], ys=ys0, src_mask=dev_in[i][1])
        barrier(world)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        row_steps = all_steps = 0
        for i in range(W, W + reps):
            sampler.decode(zs=dev_in[i][0], ys=ys0, src_mask=dev_in[i][1])
            row_steps += sampler.last_row_steps
            all_steps += sampler.last_steps_executed * BATCH
        a1.record()
        torch.cuda.synchronize()
        ams = max_over_ranks(a0.elapsed_time(a1), world, dev)
        sampler.sample_smiles(BATCH, zs=inputs[0][1], toklen=list(inputs[0][0]))
        barrier(world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(W, W + reps):
            sampler.sample_smiles(BATCH, zs=inputs[i][1], toklen=list(inputs[i][0]))
        torch.cuda.synchronize()
        at = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
        sampler.skip_finished = False
        active = {"value": world * reps * BATCH / (ams / 1e3), "e2e": world * reps * BATCH / (at / 1e3), "unit": "SMILES/s", "calls": reps,
                  "ms_per_call": ams / reps, "row_steps_fraction": row_steps / max(1, all_steps),
                  "compact_every": sampler.compact_every,
                  "note": "opt-in sampler kwarg skip_finished=True: same strings (tests/test_gpu_active_rows.py), rows x steps the step "
                          "kernels ran on / rows x steps of the plain loop = row_steps_fraction; attention of finished rows inside a "
                          "chunk is skipped as well"}
    h2d = int(np.mean([z.numel() * 4 + BATCH * z.size(1) + BATCH * 8 for _, z in inputs[W:]]))
    d2h = BATCH * MAX_STRLEN * 2          # int16 token ids
    cfg = sampler.model._cfg()
    per_step = L.lib().gct_decode_launches_per_step_at(cfg, BATCH)
    launches = K * (L.lib().gct_decode_begin_launches(cfg, int(inputs[W][1].size(1))) + (n_steps_run // max(K, 1)) * per_step)
    Sm_mean = float(np.mean([z.size(1) for _, z in inputs[W:]]))
    Sm_true = float(np.mean([np.mean(tl) for tl, _ in inputs[W:]]))       # keys actually attended (rows differ in length)
    return dict(value=value, ms_per_step=ms / K, e2e_value=world * K * BATCH / (t_e2e / 1e3), h2d=h2d, d2h=d2h, clocks=ck,
                launches=launches, sampler=sampler, Sm_mean=Sm_mean, Sm_true=Sm_true, decode_steps=n_steps_run / K, public=public,
                active=active)


def make_train_batch(B, S, nc, scaffold, seed, dev=None, pinned=False):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(int(S * 0.6), S + 1, (B,), generator=g)
    lens[0] = S
    toks = torch.randint(5, VOCAB, (B, S), generator=g)
    if scaffold:
        toks[:, scaffold] = 4
    ar = torch.arange(S)[None, :]
    src = torch.where(ar < lens[:, None], toks, torch.ones_like(toks))
    trg = torch.ones(B, S + 2, dtype=torch.long)
    trg[:, 0] = 2
    trg[:, 1:S + 1] = src
    trg[torch.arange(B), lens + 1] = 3
    batch = {"src": src, "trg": trg}
    if nc:
        batch["econds"] = torch.randn(B, nc, generator=g)
        batch["dconds"] = batch["econds"].clone()
    if pinned:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    if dev is not None:
        batch = {k: v.to(dev) for k, v in batch.items()}
    return batch


TRAIN_CASES = {
    "cfg3": dict(mt="pvaetf", S=78, sca=0, name="cfg3 pvaetf S=78 T=79"),
    "cfg4": dict(mt="pscavaetf", S=98, sca=19, name="cfg4 pscavaetf S=98 T=99 data-parallel"),
}


def run_training(args, rank, world, dev, case="cfg3", B=512, steps=None, grad_exchange="nccl"):
    """One optimiser step = masks + forward + loss + backward [+ NCCL gradient exchange] + fused Adam (FusedTrainer.step).
    `roofline`: algorithmic FLOPs (SURVEY 8d formula, fwd+bwd = 3x fwd, attention dense) / device time / sustained bf16 peak."""
    from gct_plus_b200.Model import Cvaetf
    from gct_plus_b200.Train.trainer1 import FusedTrainer
    from oracle.gct_oracle import ModelCfg, flops_forward
    torch.manual_seed(0)
    c = TRAIN_CASES[case]
    mt, S, sca, nc = c["mt"], c["S"], c["sca"], 3
    model = Cvaetf(VOCAB, VOCAB, dropout=0.1, nconds=nc, use_cond2lat=True, compute_dtype="bf16", **ARCH).to(dev).train()
    tr = FusedTrainer(model, mt, pad_id=1, lr=1e-4, warmup=8000, grad_exchange=grad_exchange)
    K, W = steps or min(args.steps, args.train_steps), args.warmup
    host = [make_train_batch(B, S, nc, sca, 50 + rank * 1000 + i, pinned=True) for i in range(4)]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
    tokens = [int((b["trg"][:, 1:] != 1).sum()) for b in host]
    for i in range(W):
        tr.step(devb[i % 4], 0.5)
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ntok = 0
    for i in range(K):
        tr.step(devb[i % 4], 0.5)
        ntok += tokens[i % 4]
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    # e2e: pinned host batch -> device each step, loss read back each step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        b = {k: v.to(dev, non_blocking=True) for k, v in host[i % 4].items()}
        tr.step(b, 0.5)
        tr.read_losses()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    cfg = ModelCfg(model_type=mt, src_vocab=VOCAB, trg_vocab=VOCAB, nconds=nc, use_cond2lat=True)
    flops = 3.0 * flops_forward(cfg, B, S, S + 1)
    _, tf_peak, src = peaks()
    loss = tr.read_losses()[0]
    achieved = flops / (ms / K / 1e3) / 1e12
    out = {"workload": f"{c['name']}, B={B}/GPU, {world} GPU(s)", "tokens_per_sec": world * ntok / (ms / 1e3),
           "padded_tokens_per_sec": world * K * B * (S + 1) / (ms / 1e3), "ms_per_step": ms / K, "steps": K,
           "e2e_tokens_per_sec": world * ntok / (t_e2e / 1e3),
           "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host[0].values())), "d2h_bytes_per_step": 16,
           "algorithmic_tflop_per_step": flops / 1e12,
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                        "traffic": None, "peak_source": src, "what": "whole optimiser step (all kernels), per GPU"},
           "tensor_frac_of_sustained_peak": achieved / tf_peak, "grad_exchange": tr.grad_exchange if world > 1 else None,
           "dtype": "bf16", "loss_finite": bool(np.isfinite(loss))}
    if tr.xchg is not None:
        tr.xchg.close()
    del tr, model
    torch.cuda.empty_cache()
    return out


def run_dp_parity(rank, world, dev, per_rank=64, S=98):
    """SURVEY 4 'N-GPU vs 1-GPU gradient equality at fixed global batch': every rank takes its shard of one global batch
    (fp32 tier, dropout off, the same eps rows), the overlapped NCCL exchange sums the shard gradients, and the result / N
    is compared with rank 0 running the WHOLE global batch alone; also the post-Adam parameters must be identical on all
    ranks.  train1.py:111-112 semantics (DDP mean of per-rank sum-loss gradients)."""
    from gct_plus_b200.Model import Cvaetf
    from gct_plus_b200.Train.trainer1 import FusedTrainer
    nc = 3
    glob = make_train_batch(per_rank * world, S, nc, 19, seed=777)
    eps = torch.randn(per_rank * world, nc + S, ARCH["latent_dim"], generator=torch.Generator().manual_seed(778))
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    res = {}
    for mode in ("overlap", "nccl"):
        torch.manual_seed(0)
        m = Cvaetf(VOCAB, VOCAB, dropout=0.0, nconds=nc, use_cond2lat=True, compute_dtype="fp32", **ARCH).to(dev).train()
        tr = FusedTrainer(m, "pscavaetf", pad_id=1, grad_exchange=mode)
        # every shard is padded to the global batch's S so that shapes (and the eps rows) line up with the 1-rank run
        tr.step({k: v[lo:hi].to(dev) for k, v in glob.items()}, 0.5, eps_noise=eps[lo:hi].to(dev))
        g = tr.grads.clone() / world
        p = m._flat.clone()
        ref = p.clone()
        torch.distributed.broadcast(ref, src=0)
        res[mode] = dict(g=g, param_diff=float((p - ref).abs().max()))
        tr.xchg.close()
        del tr, m
    torch.manual_seed(0)
    m = Cvaetf(VOCAB, VOCAB, dropout=0.0, nconds=nc, use_cond2lat=True, compute_dtype="fp32", **ARCH).to(dev).train()
    out = {"global_batch": per_rank * world, "tier": "fp32, dropout off", "workload": f"pscavaetf S={S}"}
    if rank == 0:
        # 1-rank run of the whole global batch (no exchange): gradient of the global sum-loss, then / world like DDP's mean
        tr = FusedTrainer(m, "pscavaetf", pad_id=1, grad_exchange="none")
        tr.step({k: v.to(dev) for k, v in glob.items()}, 0.5, eps_noise=eps.to(dev))
        g1 = tr.grads / world
        scale = float(g1.abs().max())
        for mode in ("overlap", "nccl"):
            out[f"grad_max_rel_diff_vs_1rank_{mode}"] = float((res[mode]["g"] - g1).abs().max()) / scale
    for mode in ("overlap", "nccl"):
        t = torch.tensor([res[mode]["param_diff"]], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out[f"param_max_abs_diff_across_ranks_{mode}"] = float(t.item())
    torch.distributed.barrier()
    del m
    torch.cuda.empty_cache()
    return out


def run_cfg5(args, rank, world, dev):
    """BASELINE cfg 5: scavaetf scaffold-conditioned sampling, 20-token scaffold (ys starts at length 22, latent length
    21 + toklen), multinomial, max_strlen 100, independent draws per rank (rank-offset seeds, no communication)."""
    from gct_plus_b200.Train.dp import rank_seed
    sampler = build_sampler(dev, model_type="scavaetf", latent_bucket=16)
    per_call = min(args.cfg5_draws, 31250)
    calls = max(1, args.cfg5_draws // per_call)
    np.random.seed(rank_seed(2024, rank) % (2 ** 31))
    g = torch.Generator().manual_seed(rank_seed(2024, rank))
    sca = len(SCAFFOLD20) + 1
    ins = []
    for _ in range(2):
        toklen = sampler.sample_toklen(per_call)
        zs = torch.randn(per_call, sca + int(max(toklen)), ARCH["latent_dim"], generator=g).pin_memory()
        ins.append((toklen, zs))
    sampler.sample_smiles(per_call, SCAFFOLD20, zs=ins[0][1], toklen=list(ins[0][0]))
    sampler.sample_smiles(per_call, SCAFFOLD20, zs=ins[1][1], toklen=list(ins[1][0]))      # second call captures the CUDA graphs
    # device-resident decode
    dev_in = []
    for toklen, zs in ins:
        Lz = zs.size(1)
        mask = torch.arange(Lz).expand(per_call, 1, Lz) < (torch.LongTensor(np.asarray(toklen)).view(per_call, 1, 1) + sca)
        dev_in.append((zs.to(dev), mask.to(dev)))
    sca_ids = [sampler.TRG.vocab.stoi[t] for t in sampler.TRG.tokenize(SCAFFOLD20)]
    ys0 = sampler.init_y(per_call, add_sos=True, sca_ids=sca_ids, add_sep=True).to(dev)
    sampler.decode(zs=dev_in[0][0], ys=ys0, src_mask=dev_in[0][1])
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(calls):
        sampler.decode(zs=dev_in[i % 2][0], ys=ys0, src_mask=dev_in[i % 2][1])
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    barrier(world)
    t0 = time.perf_counter()
    n_out = 0
    for i in range(calls):
        n_out += len(sampler.sample_smiles(per_call, SCAFFOLD20, zs=ins[i % 2][1], toklen=list(ins[i % 2][0]))[0])
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    cfg = sampler.model._cfg()
    import gct_plus_b200._lib as L
    zmode = bool(L.lib().gct_decode_begin_launches(cfg, 80) != 3 + 2 * cfg.n_layers)
    out = {"workload": f"cfg5 scavaetf, 20-token scaffold prefix, {per_call * calls} draws per rank in calls of {per_call}, max_strlen 100",
           "value": world * calls * per_call / (ms / 1e3), "unit": "SMILES/s", "ms_per_call": ms / calls,
           "e2e": world * n_out / (t_e2e / 1e3), "latent_space_cross_attention": zmode,
           "projected_s_for_1M_molecules_on_8_gpus": 125000 / (calls * per_call / (t_e2e / 1e3)) if world == 8 else None}
    del sampler
    torch.cuda.empty_cache()
    return out


def run_gpu_eager_baseline(dev, budget_rows=512):
    """BASELINE.md section 3's like-for-like incumbent: the reference's arithmetic (oracle port = plain torch ops over the
    state_dict, fp32) run eagerly on the SAME B200, TF32 off and on: (a) the un-cached multinomial Sampling.decode loop at the
    reference driver's batch 512, max_strlen 100; (b) one cfg 3 forward + loss + backward (no optimiser step)."""
    O, sd, cfg = cpu_sampler_setup()
    sd = {k: v.to(dev) for k, v in sd.items()}
    out = {"kind": "oracle port of the reference on the GPU (eager torch, fp32)", "batch": budget_rows}
    rng = np.random.RandomState(3)
    toklen = np.clip(np.rint(rng.normal(35, 7, size=budget_rows)), 13, 55).astype(int)
    Lz = int(toklen.max())
    zs = torch.randn(budget_rows, Lz, ARCH["latent_dim"], generator=torch.Generator().manual_seed(3)).to(dev)
    mask = (torch.arange(Lz).expand(budget_rows, 1, Lz) < torch.LongTensor(toklen).view(budget_rows, 1, 1)).to(dev)
    ys0 = torch.full((budget_rows, 1), 2, dtype=torch.long, device=dev)
    from gct_plus_b200.Model import Cvaetf
    torch.manual_seed(0)
    sd3 = {k: v.detach().clone().to(dev) for k, v in Cvaetf(VOCAB, VOCAB, dropout=0.1, nconds=3, use_cond2lat=True, **ARCH).state_dict().items()}
    cfg3 = O.ModelCfg(model_type="pvaetf", src_vocab=VOCAB, trg_vocab=VOCAB, nconds=3, use_cond2lat=True)
    batch = make_train_batch(512, 78, 3, 0, 50, dev=dev)
    ntok = int((batch["trg"][:, 1:] != 1).sum())
    keep = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            tag = "tf32" if tf32 else "fp32"
            with torch.no_grad():
                O.sampling_decode(sd, cfg, zs[:64], ys0[:64], mask[:64], max_strlen=12, algo="multinomial")      # warm-up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ys = O.sampling_decode(sd, cfg, zs, ys0, mask, max_strlen=MAX_STRLEN, algo="multinomial")
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            out[f"decode_smiles_per_sec_{tag}"] = budget_rows / dt
            out[f"decode_steps_{tag}"] = int(ys.size(1) - 1)
            params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd3.items()}
            eps = torch.randn(512, 81, ARCH["latent_dim"], device=dev)
            for it in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                prop, mol, mu, lv, _ = O.forward_propagation(params, cfg3, batch, 1, eps)
                loss = O.loss_function(0.5, prop, mol, None, batch["trg"][:, 1:].reshape(-1), mu, lv, False, 1)[0]
                loss.backward()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                for p_ in params.values():
                    p_.grad = None
            out[f"cfg3_fwd_bwd_ms_{tag}"] = dt * 1e3
            out[f"cfg3_fwd_bwd_tokens_per_sec_{tag}"] = ntok / dt
            del params
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = keep
    torch.cuda.empty_cache()
    return out


def run_input_pipeline(dev, batch=512, nrows=20000, iters=50):
    """batches of the training dict assembled on the device (gct_collate over the pre-tokenised corpus
    in HBM) next to the CPU oracle port of the reference's SmilesDataset + collate_fn + Field.process path, both for
    pscavaetf rows (scaffold <sep> smiles + 3 properties) of MOSES-like length.  Device leg: CUDA events; the index
    upload (B x 8 bytes) is inside the timed region."""
    import pandas as pd
    from gct_plus_b200.Utils.dataset import DeviceDataLoader, TokenisedCorpus
    from oracle import collate_oracle as CO          # cpu_baseline leg only
    atoms = ["C", "c", "N", "n", "O", "o", "S", "s", "F", "Cl", "Br", "(", ")", "[nH]", "=", "#", "1", "2", "3", "-", "[C@@H]"]
    rng = np.random.RandomState(5)
    mk = lambda lo, hi: "".join(rng.choice(atoms, size=rng.randint(lo, hi)))      # noqa: E731
    props = ["logP", "tPSA", "QED"]
    df = {"src": [mk(20, 56) for _ in range(nrows)], "src_scaffold": [mk(8, 24) for _ in range(nrows)]}
    for p_ in props:
        df[f"src_{p_}"] = rng.randn(nrows).astype(np.float32)
        df[f"trg_{p_}"] = rng.randn(nrows).astype(np.float32)
    df = pd.DataFrame(df)
    import re
    rx = re.compile(r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|\\|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])")

    class _RxField(_Field):                  # the device leg's own duck-typed fields (regex of Utils/field.py:16)
        def __init__(self, itos):
            self.vocab = _Vocab()
            self.vocab.itos = list(itos)
            self.vocab.stoi = {t: i for i, t in enumerate(self.vocab.itos)}

        def tokenize(self, smi):
            return rx.findall(smi)

    src_itos = ["<unk>", "<pad>", "<sep>"] + atoms
    trg_itos = ["<unk>", "<pad>", "<sos>", "<eos>", "<sep>"] + atoms
    t0 = time.perf_counter()
    corpus = TokenisedCorpus(df, props, _RxField(src_itos), _RxField(trg_itos), use_scaffold=True).to(dev)
    t_tok = time.perf_counter() - t0
    order = torch.randperm(nrows, generator=torch.Generator().manual_seed(0)).numpy()
    dl = DeviceDataLoader(corpus, "pscavaetf", batch, order)
    it = iter(dl)
    for _ in range(3):
        next(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nb, out_bytes = 0, 0
    for b in it:
        nb += 1
        out_bytes += sum(v.numel() * v.element_size() for v in b.values())
        if nb == iters:
            break
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nb
    # CPU leg: the oracle port with its own (torchtext-like) fields over the same vocabulary
    SRC, TRG = CO.smiles_fields(atoms, True)
    assert SRC.vocab.itos == src_itos and TRG.vocab.itos == trg_itos
    t0 = time.perf_counter()
    ncpu = 0
    for _ in CO.batches(df, order[:3 * batch], batch, "pscavaetf", SRC, TRG, props, True):
        ncpu += 1
    t_cpu = (time.perf_counter() - t0) / ncpu
    return {"workload": f"pscavaetf training batches of {batch} rows from a {nrows}-row synthetic corpus", "device_batches_per_sec": 1e3 / ms,
            "device_us_per_batch": ms * 1e3, "device_output_bytes_per_batch": out_bytes // nb,
            "one_off_tokenisation_s": t_tok, "cpu_port_batches_per_sec": 1.0 / t_cpu,
            "cpu_port": "oracle port of SmilesDataset.__getitem__ + scavaetf_collate_fn + Field.process, 1 thread (the reference uses num_workers=0)"}


def base_config(batch, world, note):
    return {"workload": "cfg2 vaetf unconditioned sampling: multinomial decode to max_strlen 100 (99 steps per batch), latent lengths "
                        "round(N(35,7^2)) in [13,55], random-init weights (no <eos> early stop)",
            "batch": batch, "d_model": 512, "layers": "6+6", "heads": 8, "latent": 128, "vocab": VOCAB, "l2": note,
            "sharding": f"{world} independent rank(s), no data-path collective"}


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(v, world, dev):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# =============================================================================================
# CPU arm: the oracle port of the reference's un-cached Sampling.decode on the host cores
# =============================================================================================
def cpu_sampler_setup():
    from oracle import gct_oracle as O
    from gct_plus_b200.Model import Vaetf          # only to reproduce the reference's init (no compute on this path)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in Vaetf(VOCAB, VOCAB, dropout=0.1, nconds=0, **ARCH).state_dict().items()}
    cfg = O.ModelCfg(model_type="vaetf", src_vocab=VOCAB, trg_vocab=VOCAB)
    return O, sd, cfg


def cpu_decode(O, sd, cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    toklen = np.clip(np.rint(rng.normal(35, 7, size=n)), 13, 55).astype(int)
    Lz = int(toklen.max())
    zs = torch.randn(n, Lz, ARCH["latent_dim"], generator=g)
    mask = torch.arange(Lz).expand(n, 1, Lz) < torch.LongTensor(toklen).view(n, 1, 1)
    ys0 = torch.full((n, 1), 2, dtype=torch.long)
    t0 = time.perf_counter()
    with torch.no_grad():
        ys = O.sampling_decode(sd, cfg, zs, ys0, mask, max_strlen=MAX_STRLEN, algo="multinomial")
    dt = time.perf_counter() - t0
    [O.id_to_smi(r.tolist(), ITOS) for r in ys]
    return dt, ys.size(1) - 1


def cpu_baseline(budget_s=20.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, sd, cfg = cpu_sampler_setup()
    dt1, _ = cpu_decode(O, sd, cfg, 2, 1)               # calibration + warm-up
    n = int(max(2, min(64, budget_s / max(dt1 / 2, 1e-3) * 0.6)))
    dt, steps = cpu_decode(O, sd, cfg, n, 2)
    return {"value": n / dt, "unit": "SMILES/s", "cores": cores, "kind": "port",
            "sample": f"oracle port of the reference's un-cached multinomial Sampling.decode, one batch of {n} latents, "
                      f"{steps} steps (max_strlen {MAX_STRLEN}), fp32, torch {torch.__version__} CPU with {cores} threads"}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, sd, cfg = cpu_sampler_setup()
    K, W = args.steps, args.warmup
    dt1, _ = cpu_decode(O, sd, cfg, 1, 1)
    n = int(max(1, min(32, 150.0 / (K + W) / max(dt1, 1e-3) * 0.7)))
    for i in range(W):
        cpu_decode(O, sd, cfg, n, 10 + i)
    t0 = time.perf_counter()
    for i in range(K):
        cpu_decode(O, sd, cfg, n, 100 + i)
    dt = time.perf_counter() - t0
    v = K * n / dt
    sample = (f"each step = oracle port of the reference's un-cached multinomial Sampling.decode on {n} latent draw(s), "
              f"{MAX_STRLEN - 1} steps, fp32 CPU, {cores} threads")
    print(json.dumps({"impl": "reference", "metric": "sampled_smiles_per_sec", "value": v, "unit": "SMILES/s", "n_gpus": args.gpus,
                      "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": base_config(n, 1, f"CPU oracle port; each step decodes {n} draw(s) instead of {BATCH}"),
                      "cpu_baseline": {"value": v, "unit": "SMILES/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": "SMILES/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)        # 8 sample_smiles calls of --batch draws
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=30000, help="latent draws per sample_smiles call (cfg 2: 30k draws = one call)")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--train-batch", type=int, default=512)
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batch512", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-public-call", dest="public_call", action="store_false")
    ap.add_argument("--cfg5-draws", type=int, default=125000, help="scaffold-conditioned draws per rank (cfg 5: 1M over 8 GPUs)")
    args = ap.parse_args()
    global BATCH
    BATCH = args.batch
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (gct_plus_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    hbm_peak, tf_peak, peak_src = peaks()
    s = run_sampling(args, rank, world, dev)
    b512 = None
    if BATCH != 512 and not args.no_batch512:
        big = BATCH
        BATCH = 512
        a2 = argparse.Namespace(**vars(args))
        a2.steps, a2.warmup, a2.public_call, a2.active_rows = 12, 3, False, False
        r = run_sampling(a2, rank, world, dev)
        b512 = {"batch": 512, "value": r["value"], "unit": "SMILES/s", "ms_per_step": r["ms_per_step"], "steps": 12,
                "e2e": r["e2e_value"], "note": "the reference driver's default -batch_size (uc_sampling.py:16-23)"}
        BATCH = big
    steps = MAX_STRLEN - 1
    alg_per_launch, ms_per_launch, nl, roof_rows = time_decode_attention(s["sampler"], dev, steps, s["Sm_mean"])
    achieved = alg_per_launch / (ms_per_launch / 1e3) / 1e9
    total_alg = decode_alg_bytes(6, BATCH, s["Sm_true"], steps, 2, latent_form=True)
    total_kv_form = decode_alg_bytes(6, BATCH, s["Sm_true"], steps, 2)
    s.pop("sampler")
    torch.cuda.empty_cache()
    cfg5 = None if args.no_cfg5 else run_cfg5(args, rank, world, dev)
    train, dp_parity, eager = None, None, None
    if not args.no_train:
        train = {}
        if world == 1:
            train["cfg3"] = run_training(args, rank, world, dev, "cfg3", args.train_batch)
        train["cfg4"] = run_training(args, rank, world, dev, "cfg4", args.train_batch)
        train["cfg4_batch64"] = run_training(args, rank, world, dev, "cfg4", 64, steps=min(args.train_steps, 30))
        if world > 1:
            train["cfg4_overlapped_exchange"] = run_training(args, rank, world, dev, "cfg4", args.train_batch, grad_exchange="overlap")
            dp_parity = run_dp_parity(rank, world, dev)
        if rank == 0 and not args.no_cpu:
            train["input_pipeline"] = run_input_pipeline(dev)
    if rank == 0 and not args.no_eager and world == 1:
        eager = run_gpu_eager_baseline(dev)
    if rank == 0:
        cpu = None if args.no_cpu else cpu_baseline()
        line = {"metric": "sampled_smiles_per_sec", "value": s["value"], "unit": "SMILES/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": base_config(BATCH, world, f"working set (KV cache {0.9 * BATCH / 512:.1f} GB per batch) larger than L2, no flush needed"),
                "e2e": {"value": s["e2e_value"], "unit": "SMILES/s", "h2d_bytes_per_step": s["h2d"], "d2h_bytes_per_step": s["d2h"],
                        "call": "sampler.sample_smiles(n, zs=<pinned host latents>, toklen=<host list>) -> Python strings"},
                "e2e_public_call": s["public"], "active_row_decode": s["active"],
                "gpu_launches": s["launches"], "clocks": s["clocks"],
                "roofline": {"bound": "hbm", "kernel": "damma::decode_attn_mma_kernel<2> (self-attention over the KV cache, bf16)",
                             "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": NCU_TRAFFIC["dram_bytes_per_launch"], "traffic_detail": NCU_TRAFFIC,
                             "peak_source": peak_src, "launches_timed": nl, "us_per_launch": ms_per_launch * 1e3,
                             "rows_per_launch": roof_rows,
                             "whole_decode": {"algorithmic_bytes_per_batch": total_alg,
                                              "achieved_GBps": total_alg / (s["ms_per_step"] / 1e3) / 1e9,
                                              "frac": total_alg / (s["ms_per_step"] / 1e3) / 1e9 / hbm_peak,
                                              "note": "bytes of this build's algorithm (cross-attention over the latent rows); "
                                                      "kv_form_* = SURVEY 8(d)'s K/V-projection form of the same decode",
                                              "kv_form_bytes_per_batch": total_kv_form,
                                              "kv_form_equivalent_GBps": total_kv_form / (s["ms_per_step"] / 1e3) / 1e9}},
                "cpu_baseline": cpu, "gpu_eager_baseline": eager, "batch512": b512, "cfg5": cfg5, "train": train, "dp_parity": dp_parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
