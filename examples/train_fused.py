"""train1.py-style data-parallel training on the device path, end to end, on a synthetic corpus:

    python examples/train_fused.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_fused.py

DataloaderPreparation (pre-tokenised corpus in HBM, gct_collate batches) -> FusedTrainer (fwd + loss + bwd + NCCL allreduce +
fused Adam + Noam LR) -> save_checkpoint in the reference's layout.  Replace `synthetic_frame` / `smiles_fields` by the
reference's dataframe and pickled torchtext Fields (Utils/field.py:47-63) for real data.
"""
import argparse
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gct_plus_b200.Model.build_model import get_model  # noqa: E402
from gct_plus_b200.Train.trainer1 import FusedTrainer, KLAnnealer, save_checkpoint  # noqa: E402
from gct_plus_b200.Utils.dataset import DataloaderPreparation  # noqa: E402

ATOMS = ["C", "c", "N", "n", "O", "o", "S", "s", "F", "Cl", "Br", "(", ")", "[nH]", "=", "#", "1", "2", "3", "-", "[C@@H]"]
PROPS = ["logP", "tPSA", "QED"]


class _Vocab:
    def __init__(self, itos):
        self.itos = list(itos)
        self.stoi = {t: i for i, t in enumerate(self.itos)}

    def __len__(self):
        return len(self.itos)


class _Field:                      # duck-typed like torchtext's Field (only what the path touches)
    batch_first = True

    def __init__(self, itos):
        import re
        self.vocab = _Vocab(itos)
        self._re = re.compile(r"(\[[^\]]+]|Br?|Cl?|N|O|S|P|F|I|b|c|n|o|s|p|\(|\)|\.|=|#|-|\+|\\|\/|:|~|@|\?|>|\*|\$|\%[0-9]{2}|[0-9])")

    def tokenize(self, s):
        return self._re.findall(s)


def synthetic_frame(n, seed):
    rng = np.random.RandomState(seed)
    mk = lambda lo, hi: "".join(rng.choice(ATOMS, size=rng.randint(lo, hi)))      # noqa: E731
    d = {"src": [mk(20, 56) for _ in range(n)], "src_scaffold": [mk(6, 20) for _ in range(n)]}
    for p in PROPS:
        d[f"src_{p}"] = rng.randn(n).astype(np.float32)
        d[f"trg_{p}"] = d[f"src_{p}"]
    return pd.DataFrame(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--out", default="/tmp/gct_example")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    SRC = _Field(["<unk>", "<pad>", "<sep>"] + ATOMS)
    TRG = _Field(["<unk>", "<pad>", "<sos>", "<eos>", "<sep>"] + ATOMS)
    args = argparse.Namespace(N=6, d_model=512, d_ff=2048, H=8, latent_dim=128, dropout=0.1, use_cond2dec=False, use_cond2lat=True,
                              variational=True, property_list=PROPS, get_attn=False, model_type="pscavaetf", pad_id=1)
    torch.manual_seed(0)
    model = get_model(args, len(SRC.vocab), len(TRG.vocab), local).to(torch.device("cuda", local)).train()
    trainer = FusedTrainer(model, args.model_type, pad_id=args.pad_id, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, warmup=8000)
    prep = DataloaderPreparation(local, SRC, TRG, args.model_type, PROPS, world_size=world, use_scaffold=True)
    if world > 1:
        prep.rank = local
    loader = prep.get_dataloader(synthetic_frame(a.rows, seed=1), a.batch, is_train=True)
    if world > 1:      # DistributedSampler wants the global rank
        from torch.utils.data import DistributedSampler
        from gct_plus_b200.Utils.dataset import _Rows
        loader.sampler = DistributedSampler(_Rows(a.rows), world, rank, shuffle=True)
    os.makedirs(a.out, exist_ok=True)
    for epoch in range(1, a.epochs + 1):
        if world > 1:
            loader.sampler.set_epoch(epoch)
        beta = KLAnnealer(epoch, 0.02, 0.02, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ntok = 0
        for batch in loader:
            trainer.step(batch, beta)
            ntok += int(batch["trg"].numel())
        e1.record()
        torch.cuda.synchronize()
        loss, rce, kld = trainer.read_losses()
        if rank == 0:
            print(f"epoch {epoch}: last batch loss/row {loss / a.batch:.3f} (RCE {rce / a.batch:.3f}, KLD {kld / a.batch:.3f}), "
                  f"{world * ntok / (e0.elapsed_time(e1) / 1e3) / 1e6:.2f} M padded target tokens/s", flush=True)
            save_checkpoint(args, model, trainer, os.path.join(a.out, f"model_{epoch}.pt"))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
