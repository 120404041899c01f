"""world_size-2 gloo checks of the data-parallel host logic (no GPU): gradient exchange semantics of
train1.py's DDP (mean over ranks of per-rank sum-loss gradients) and the sampling shard arithmetic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, cfg_from_fixture, load_golden, sd_from_fixture


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(sd, cfg, batch, eps, beta):
    params = {k: v.clone().requires_grad_(not k.endswith("pe.pe")) for k, v in sd.items()}
    prop, mol, mu, lv, _ = O.forward_propagation(params, cfg, batch, 1, eps)
    loss = O.loss_function(beta, prop, mol, None, batch["trg"][:, 1:].reshape(-1), mu, lv, False, 1)[0]
    loss.backward()
    keys = [k for k, p in params.items() if p.grad is not None]
    return keys, torch.cat([params[k].grad.flatten() for k in keys])


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gct_plus_b200.Train.dp import allreduce_sum_, rank_seed, shard_range, world_info
    assert world_info() == (rank, world)
    fx = load_golden("scavaetf_small")
    cfg, sd = cfg_from_fixture(fx), sd_from_fixture(fx)
    B = fx["batch"]["src"].size(0)
    torch.manual_seed(3)
    eps = torch.randn(B, fx["batch"]["src"].size(1), cfg.latent_dim)
    lo, hi = shard_range(B, rank, world)
    shard = {k: v[lo:hi] for k, v in fx["batch"].items()}
    _, flat = _grads(sd, cfg, shard, eps[lo:hi], 0.3)
    scale = allreduce_sum_(flat)
    if rank == 0:
        _, full = _grads(sd, cfg, fx["batch"], eps, 0.3)
        # DDP: mean over ranks of per-rank sum-loss gradients == full-batch sum-loss gradient / world
        err = float((flat * scale - full / world).abs().max() / full.abs().max())
        out.put((err, scale, rank_seed(5, 0) != rank_seed(5, 1)))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradient_mean_semantics_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale, seeds_differ = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert scale == 0.5 and seeds_differ
    assert err < 1e-5, err


def test_shard_range_partitions_everything():
    from gct_plus_b200.Train.dp import shard_range
    for n in (0, 1, 7, 30000, 1000000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
