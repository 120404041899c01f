"""Drop-in for Train/trainer1.py (loss_function, KLAnnealer, save_checkpoint, run_epoch,
train_model) plus ``FusedTrainer``: the same optimisation step with no autograd and no
per-step host syncs (forward -> loss -> backward -> [NCCL allreduce] -> fused Adam).
"""
from __future__ import annotations

import ctypes as C
import os
from functools import reduce
from time import time

import torch
import torch.distributed as dist

from .. import _lib as L
from ..Model.forward_propagation1 import forward_propagation
from ..Model.modules import get_src_mask, get_trg_mask
from .dp import GradExchange, allreduce_sum_, world_info


def KLAnnealer(epoch, KLA_ini_beta, KLA_inc_beta, KLA_beg_epoch):
    return KLA_ini_beta + KLA_inc_beta * ((epoch + 1) - KLA_beg_epoch)


def noam_lr(step, d_model, warmup):
    """d^-0.5 * min(step^-0.5, step * warmup^-1.5)  (reference trainer1.py:117-123)."""
    return L.lib().gct_noam_lr(int(step), int(d_model), int(warmup))


class _LossFn(torch.autograd.Function):
    """CE(sum, ignore pad) + beta * KL(sum) in one pass; gradients are produced by the same kernels."""

    @staticmethod
    def forward(ctx, preds_mol, mu, log_var, ys_mol, beta, pad_id):
        lib = L.lib()
        L.require_cuda(preds_mol, "preds_mol")
        V = preds_mol.size(-1)
        logits = preds_mol.float().contiguous().view(-1, V)
        mu_c, lv_c = mu.float().contiguous(), log_var.float().contiguous()
        rows, nlat = logits.size(0), mu_c.numel()
        tgt = ys_mol.contiguous().view(-1)
        out4 = torch.empty(4, device=logits.device, dtype=torch.float32)
        dlogits = torch.empty_like(logits)
        dmu, dlv = torch.empty_like(mu_c), torch.empty_like(lv_c)
        scratch = torch.empty(lib.gct_loss_scratch_bytes(rows, nlat), dtype=torch.uint8, device=logits.device)
        L.check(lib.gct_loss_fwd_bwd(L.ptr(logits), V, V, L.ptr(tgt), rows, int(pad_id), L.ptr(mu_c), L.ptr(lv_c), nlat,
                                     float(beta), 1.0, L.ptr(out4), L.ptr(dlogits), L.ptr(dmu), L.ptr(dlv), L.ptr(scratch),
                                     L.stream_ptr()), "gct_loss_fwd_bwd")
        ctx.save_for_backward(dlogits.view(preds_mol.shape), dmu.view(mu.shape), dlv.view(log_var.shape))
        loss, rce, kld = out4[0], out4[1], out4[3]
        ctx.mark_non_differentiable(rce, kld)
        return loss, rce, kld

    @staticmethod
    def backward(ctx, g, _g1, _g2):
        dlogits, dmu, dlv = ctx.saved_tensors
        return dlogits * g, dmu * g, dlv * g, None, None, None


def loss_function(beta, preds_prop, preds_mol, ys_cond, ys_mol, mu, log_var, use_cond2dec, pad_id):
    """Same signature and return tuple as the reference: (loss, RCE_mol, RCE_prop, KLD)."""
    loss, rce, kld = _LossFn.apply(preds_mol, mu, log_var, ys_mol, beta, pad_id)
    if use_cond2dec:
        # property head only exists with -use_cond2dec, which no shipped script sets
        rce_prop = torch.nn.functional.mse_loss(preds_prop, ys_cond, reduction='sum')
        loss = loss + rce_prop
    else:
        rce_prop = torch.zeros(1)
    return loss, rce, rce_prop, kld


def save_checkpoint(args, model, optimizer, save_path):
    names = ('N', 'd_model', 'd_ff', 'H', 'latent_dim', 'dropout', 'use_cond2dec', 'use_cond2lat', 'variational')
    params = {'nconds': len(args.property_list)}
    for name in names:
        params[name] = getattr(args, name)
    torch.save({'model_state_dict': model.state_dict(), 'opt_state_dict': optimizer.state_dict(),
                'model_params': params}, save_path)


def run_epoch(args, model, optimizer, dataloader, current_step, beta, LOG, train):
    """One pass over `dataloader`, reference semantics (trainer1.py:71-157): LR is written into the
    optimiser AFTER the step, three scalar reads per step, a log line per step."""
    history = {'RCE': [], 'KLD': [], 'LOSS': [], 'BETA': [], 'LR': []}
    model_cost_time = update_cost_time = 0
    cost_time = -time()
    lr = None
    for i, batch in enumerate(dataloader):
        current_step += 1
        n_onebatch = batch['src'].size(0)
        model_cost_time -= time()
        preds_prop, preds_mol, mu, log_var, _ = forward_propagation[args.model_type](
            model, batch, args.pad_id, args.use_cond2dec)
        model_cost_time += time()
        nprop = len(args.property_list)
        ys_cond = torch.unsqueeze(batch['dconds'], 2).contiguous().view(-1, nprop, 1) if nprop > 0 else None
        ys_mol = batch['trg'][:, 1:].contiguous().view(-1)
        update_cost_time -= time()
        if train:
            optimizer.zero_grad(set_to_none=True)
        loss, RCE_mol, RCE_prop, KLD = loss_function(beta, preds_prop, preds_mol, ys_cond, ys_mol, mu, log_var,
                                                     args.use_cond2dec, args.pad_id)
        if train:
            loss.backward()
            optimizer.step()
        update_cost_time += time()
        if args.lr_scheduler == "WarmUpDefault":
            lr = noam_lr(current_step, args.d_model, args.lr_WarmUpSteps)
        if train and lr is not None:
            for group in optimizer.param_groups:
                group['lr'] = lr
        current_lr = optimizer.param_groups[-1]['lr']
        history['RCE'].append(RCE_mol.item() / n_onebatch)
        history['KLD'].append(KLD.item() / n_onebatch)
        history['LOSS'].append(loss.item() / n_onebatch)
        history['BETA'].append(beta)
        history['LR'].append(current_lr)
        if LOG is not None:
            LOG.info(f'{i + 1}/{len(dataloader):<10}\tRCE: {history["RCE"][-1]:.5f}\tKLD: {history["KLD"][-1]:.5f}\t'
                     f'LOSS: {history["LOSS"][-1]:.5f}\tTIME(s): {time() + cost_time:.1f}\t'
                     f'MODELTIME(s): {model_cost_time:.1f}\tUPDATETIME(s): {update_cost_time:.1f}')
    return (history, current_step) if train else history


def train_model(args, model, optimizer, train_loader, valid_loader, rank, world_size, LOG):
    """Epoch loop of the reference (trainer1.py:159-255): KL annealing, barriers, per-rank CSVs merged
    on rank 0, checkpoint per epoch."""
    import pandas as pd
    beta = 0
    current_step = (args.start_epoch - 1) * len(train_loader)
    for epoch in range(args.start_epoch, args.num_epoch + 1):
        if world_size > 1:
            train_loader.sampler.set_epoch(epoch)
        if args.use_KLA:
            if epoch + 1 >= args.KLA_beg_epoch and beta < args.KLA_max_beta:
                beta = KLAnnealer(epoch, args.KLA_ini_beta, args.KLA_inc_beta, args.KLA_beg_epoch)
        else:
            beta = 1
        if world_size > 1:
            dist.barrier()
        model.train()
        train_history, current_step = run_epoch(args, model, optimizer, train_loader, current_step, beta, LOG, train=True)
        suffix = f'_r{rank}' if world_size > 1 else ''
        pd.DataFrame(train_history).to_csv(os.path.join(args.model_folder, f'train_{epoch}{suffix}.csv'))
        if world_size > 1:
            dist.barrier()
        model.eval()
        with torch.no_grad():
            valid_history = run_epoch(args, model, optimizer, valid_loader, current_step, beta, LOG, train=False)
        pd.DataFrame(valid_history).to_csv(os.path.join(args.model_folder, f'valid_{epoch}{suffix}.csv'))
        if world_size > 1:
            dist.barrier()
        if rank == 0:
            save_checkpoint(args, model, optimizer, os.path.join(args.model_folder, f'model_{epoch}.pt'))
        if world_size > 1 and rank == 0:
            for kind in ('train', 'valid'):
                his = [pd.read_csv(os.path.join(args.model_folder, f'{kind}_{epoch}_r{r}.csv'), index_col=[0])
                       for r in range(world_size)]
                cols = [reduce(lambda x, y: x[[c]] + y[[c]], his) / world_size for c in ('RCE', 'KLD', 'LOSS')]
                pd.concat(cols + [his[0]['BETA'], his[0]['LR']], axis=1).to_csv(
                    os.path.join(args.model_folder, f'{kind}_{epoch}.csv'))
        if world_size > 1:
            dist.barrier()


class FusedTrainer:
    """The optimisation step of train1.py / trainer1.py without autograd:

        masks -> gct_forward -> gct_loss_fwd_bwd -> gct_backward -> [allreduce(sum)] -> gct_adam_step

    Semantics kept from the reference: loss = CE_sum + beta*KL_sum per rank, gradient = mean over
    ranks of the per-rank gradient (DDP), Adam(lr, betas=(0.9, 0.98), eps=1e-9), and the Noam LR
    computed from `current_step` is applied to the NEXT step (trainer1.py:112-127).  Losses stay on
    the device; `read_losses()` does the one host read when the caller wants numbers.
    """

    def __init__(self, model, model_type, pad_id=1, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, warmup=8000,
                 use_cond2dec=False, process_group=None, grad_exchange="nccl", force_exchange=False):
        self.model, self.model_type, self.pad_id = model, model_type, pad_id
        self.lr, self.betas, self.eps, self.warmup = lr, betas, eps, warmup
        # use_cond2dec (cvaetf.py:103-105,184-186; trainer1.py:24-26): nconds property rows precede the target rows of the
        # decoder; they carry no cross-entropy term, the prop_fc head on them adds MSE(sum) to the loss (gct_prop_head_fwd_bwd)
        self.use_cond2dec = bool(getattr(model, 'use_cond2dec', False) and int(model.nconds) > 0)
        if use_cond2dec and not self.use_cond2dec:
            raise L.GctError("use_cond2dec=True but the model was not built with use_cond2dec / nconds > 0")
        if self.use_cond2dec and model_type not in ('pvaetf', 'pscavaetf'):
            raise L.GctError("use_cond2dec needs a property-conditioned model type (pvaetf / pscavaetf)")
        self.variational = model._variational()
        self.pg = process_group
        self.world = world_info(process_group)[1]
        # gradient exchange across ranks: "nccl" (default) = one all-reduce of the flat buffer after the backward through the
        # library's communicator; "overlap" = bucketed NCCL all-reduce on a side stream while the backward still runs
        # (gct_backward_dp, DDP's behaviour); "torch" = one all-reduce through torch.distributed (any backend).
        # [B200, 2 GPUs, cfg 4, profiles/r02_ab_dp_exchange_2gpu.txt] no exchange 27.92 ms, "nccl" 28.51-28.55, "overlap"
        # 28.55-28.73: the NCCL kernels do run concurrently with the weight-gradient GEMMs (profiles/r02_dp_overlap_trace_2gpu.json:
        # 100 % of their 1.0 ms), but every SM they occupy pushes CTAs of a statically scheduled persistent GEMM into a second
        # wave, which costs what the overlap hides -- so the plain form is the default.
        assert grad_exchange in ("overlap", "nccl", "torch", "none")
        if grad_exchange == "none":        # single-process semantics even inside an initialised process group (reference runs, parity checks)
            self.world, grad_exchange = 1, "torch"
        self.grad_exchange = grad_exchange if (self.world > 1 or force_exchange) else "torch"     # force_exchange: 1-rank communicator (tests)
        self.xchg = GradExchange(model, process_group) if self.grad_exchange != "torch" else None
        flat = model._flat
        self.grads = torch.zeros_like(flat)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = 0
        self.out4 = torch.zeros(4, device=flat.device, dtype=torch.float32)
        self.has_conds = model_type in ('pvaetf', 'pscavaetf')
        model.pad_id = pad_id
        self._bufs = {}

    def step(self, batch, beta, eps_noise=None, train=True):
        m, lib = self.model, L.lib()
        cfg = m._cfg()
        dev = m._flat.device
        src, trg = batch['src'].contiguous(), batch['trg']
        trg_in = trg[:, :-1].contiguous()
        B, S = src.shape
        T = trg_in.size(1)
        nc = cfg.nconds if self.has_conds else 0
        econds = batch['econds'].float().contiguous() if nc else None
        dconds = batch['dconds'].float().contiguous() if nc else None
        Se, lat = nc + S, cfg.latent_dim
        Ld = T + (nc if self.use_cond2dec else 0)
        if self.use_cond2dec:       # property rows get the ignore index: no cross-entropy term, zero dlogits from the CE kernel
            ys_mol = torch.cat([torch.full((B, nc), int(self.pad_id), dtype=trg.dtype, device=trg.device), trg[:, 1:]], dim=1).contiguous().view(-1)
        else:
            ys_mol = trg[:, 1:].contiguous().view(-1)
        key = (B, S, T)
        bf = self._bufs.get(key)
        if bf is None:
            f32 = dict(device=dev, dtype=torch.float32)
            bf = dict(sm=torch.empty((B, Se), device=dev, dtype=torch.uint8),
                      tm=torch.empty((B, Ld, Ld), device=dev, dtype=torch.uint8),
                      logits=torch.empty((B, Ld, cfg.trg_vocab), **f32), dlogits=torch.empty((B, Ld, cfg.trg_vocab), **f32),
                      mu=torch.empty((B, Se, lat), **f32), lv=torch.empty((B, Se, lat), **f32), z=torch.empty((B, Se, lat), **f32),
                      dmu=torch.empty((B, Se, lat), **f32), dlv=torch.empty((B, Se, lat), **f32),
                      eps=torch.empty((B, Se, lat), **f32),
                      ws=torch.empty(lib.gct_forward_workspace_bytes(C.byref(cfg), B, S, T), dtype=torch.uint8, device=dev),
                      scratch=torch.empty(lib.gct_backward_scratch_bytes(C.byref(cfg), B, S, T), dtype=torch.uint8, device=dev),
                      lscr=torch.empty(lib.gct_loss_scratch_bytes(B * Ld, B * Se * lat), dtype=torch.uint8, device=dev))
            self._bufs = {key: bf}          # keep one shape resident
        st = L.stream_ptr()
        L.check(lib.gct_src_mask(L.ptr(src), B, S, nc, self.pad_id, L.ptr(bf['sm']), st), "gct_src_mask")
        L.check(lib.gct_trg_mask(L.ptr(trg_in), B, T, nc if cfg.use_cond2dec else 0, self.pad_id, L.ptr(bf['tm']), st), "gct_trg_mask")
        if not self.variational:
            eps_ptr = None                      # z = mu (Sampler / Encoder.sampling with variational=False)
        else:
            if eps_noise is None:
                bf['eps'].normal_()
            else:
                bf['eps'].copy_(eps_noise)
            eps_ptr = bf['eps'].data_ptr()
        if m._step_seed is None:
            m._step_seed = torch.initial_seed() & 0xFFFFFFFF
        m._step_seed = (m._step_seed * 1664525 + 1013904223) & 0xFFFFFFFF
        w = m._weights(self.grads, trust_shadow=True)
        io = L.GctIO(src=src.data_ptr(), trg=trg_in.data_ptr(), src_mask=bf['sm'].data_ptr(), trg_mask=bf['tm'].data_ptr(),
                     econds=econds.data_ptr() if econds is not None else None,
                     dconds=dconds.data_ptr() if dconds is not None else None, eps=eps_ptr, z_in=None,
                     B=B, S=S, T=T, train=int(train), seed=m._step_seed, run_encoder=1, run_decoder=1,
                     logits=bf['logits'].data_ptr(), mu=bf['mu'].data_ptr(), log_var=bf['lv'].data_ptr(), z=bf['z'].data_ptr(),
                     enc_attn=None, dec_attn1=None, dec_attn2=None)
        L.check(lib.gct_forward(C.byref(cfg), C.byref(w), C.byref(io), L.ptr(bf['ws']), bf['ws'].numel(), st), "gct_forward")
        V = cfg.trg_vocab
        L.check(lib.gct_loss_fwd_bwd(L.ptr(bf['logits']), V, V, L.ptr(ys_mol), B * Ld, self.pad_id, L.ptr(bf['mu']), L.ptr(bf['lv']),
                                     B * Se * lat, float(beta), 1.0, L.ptr(self.out4), L.ptr(bf['dlogits']) if train else None,
                                     L.ptr(bf['dmu']) if train else None, L.ptr(bf['dlv']) if train else None, L.ptr(bf['lscr']), st),
                "gct_loss_fwd_bwd")
        if train:
            self.grads.zero_()
        if self.use_cond2dec:
            off = m._offsets
            gw, gb = int(off[18]), int(off[19])          # GCT_SLOT_PROP_W / GCT_SLOT_PROP_B
            L.check(lib.gct_prop_head_fwd_bwd(L.ptr(bf['logits']), B, Ld, nc, V, m._flat[gw:].data_ptr(), m._flat[gb:].data_ptr(),
                                              L.ptr(dconds), 1.0, None, L.ptr(self.out4), L.ptr(bf['dlogits']) if train else None,
                                              self.grads[gw:].data_ptr() if train else None, self.grads[gb:].data_ptr() if train else None, st),
                    "gct_prop_head_fwd_bwd")
        if not train:
            return self.out4
        if self.grad_exchange == "overlap":
            x = self.xchg
            L.check(lib.gct_backward_dp(C.byref(cfg), C.byref(w), C.byref(io), L.ptr(bf['dlogits']), L.ptr(bf['dmu']), L.ptr(bf['dlv']),
                                        None, L.ptr(bf['ws']), bf['ws'].numel(), L.ptr(bf['scratch']), bf['scratch'].numel(), x.comm,
                                        C.cast(x.buckets, C.c_void_p), x.n_buckets, C.c_void_p(x.stream.cuda_stream), st),
                    "gct_backward_dp")
            gscale = 1.0 / self.world
        else:
            L.check(lib.gct_backward(C.byref(cfg), C.byref(w), C.byref(io), L.ptr(bf['dlogits']), L.ptr(bf['dmu']), L.ptr(bf['dlv']),
                                     None, L.ptr(bf['ws']), bf['ws'].numel(), L.ptr(bf['scratch']), bf['scratch'].numel(), st),
                    "gct_backward")
            if self.grad_exchange == "nccl":
                self.xchg.allreduce_(self.grads)
                gscale = 1.0 / self.world
            else:
                # DDP semantics: mean over ranks of the per-rank sum-loss gradient
                gscale = allreduce_sum_(self.grads, self.pg) if self.world > 1 else 1.0
        self.step_count += 1
        shadow = m._shadow if m.compute_dtype == "bf16" else None
        L.check(lib.gct_adam_step(L.ptr(m._flat), L.ptr(self.grads), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                  L.ptr(shadow), m._flat.numel(), self.step_count, float(self.lr), self.betas[0], self.betas[1],
                                  self.eps, gscale, st), "gct_adam_step")
        m._shadow_fresh = shadow is not None       # Adam wrote master and shadow in the same pass
        self.lr = noam_lr(self.step_count, cfg.d_model, self.warmup)       # takes effect on the next step
        return self.out4

    def read_losses(self, with_prop=False):
        """(loss, RCE_mol, KLD) of the last step -- one host read; with_prop=True: (loss, RCE_mol, RCE_prop, KLD) like loss_function."""
        loss, rce, rprop, kld = self.out4.tolist()
        return (loss, rce, rprop, kld) if with_prop else (loss, rce, kld)

    # ------------------------------------------------------------------ checkpoint compatibility
    def state_dict(self):
        """The optimiser state in torch.optim.Adam's state_dict layout (parameter order = model.parameters(), as in the
        reference's Adam(model.parameters(), ...), train1.py:116-119), so `save_checkpoint(args, model, trainer, path)`
        writes the same 'opt_state_dict' a reference run would and reference tooling can load it."""
        m = self.model
        state = {}
        if self.step_count > 0:
            for i, (off, n, shape) in enumerate(m._grad_views):
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': self.exp_avg[off:off + n].view(shape).clone(),
                            'exp_avg_sq': self.exp_avg_sq[off:off + n].view(shape).clone()}
        group = {'lr': float(self.lr), 'betas': tuple(self.betas), 'eps': float(self.eps), 'weight_decay': 0, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'params': list(range(len(m._param_list)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        """Resumes from an Adam state_dict (this class's or torch.optim.Adam's / the reference's 'opt_state_dict')."""
        m = self.model
        group = sd['param_groups'][0]
        if len(group['params']) != len(m._param_list):
            raise L.GctError("optimizer state has %d parameters, the model %d" % (len(group['params']), len(m._param_list)))
        self.lr, self.betas, self.eps = float(group['lr']), tuple(group['betas']), float(group['eps'])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, p in enumerate(m._param_list):
            st = sd['state'].get(i, sd['state'].get(str(i)))
            if st is None:
                continue
            off, n, _ = m._grad_views[i]
            self.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise L.GctError("per-parameter Adam step counts differ: %s" % sorted(steps))
        self.step_count = steps.pop() if steps else 0
