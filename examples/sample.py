"""uc_sampling.py-style unconditioned sampling on the device path:  python examples/sample.py [--n 30000]

get_model -> sampling_tool_dict['vaetf'] -> sample_smiles(n) (KV-cached multinomial decode, latent-space cross-attention,
host-side batch detokeniser) -> CSV with the reference's columns (uc_sampling.py:134-137).  Random-init weights unless
--model_path points at a reference checkpoint (model_{epoch}.pt).
"""
import argparse
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
from train_fused import ATOMS, _Field  # noqa: E402
from gct_plus_b200.Inference.sampling_tool import sampling_tool_dict  # noqa: E402
from gct_plus_b200.Model.build_model import get_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30000)
    ap.add_argument("--model_path", default=None)
    ap.add_argument("--out", default="/tmp/gct_example/samples.csv")
    ap.add_argument("--z_on_device", action="store_true", help="draw the latents with the CUDA generator (not reference-seeded)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    SRC = _Field(["<unk>", "<pad>"] + ATOMS)
    TRG = _Field(["<unk>", "<pad>", "<sos>", "<eos>"] + ATOMS)
    args = argparse.Namespace(N=6, d_model=512, d_ff=2048, H=8, latent_dim=128, dropout=0.1, use_cond2dec=False, use_cond2lat=False,
                              variational=True, property_list=[], get_attn=False, model_type="vaetf", pad_id=1)
    if a.model_path:
        args.model_path = a.model_path
    torch.manual_seed(0)
    np.random.seed(0)
    model = get_model(args, len(SRC.vocab), len(TRG.vocab), 0).to(dev).eval()
    toklen_data = np.clip(np.rint(np.random.RandomState(7).normal(35, 7, size=20000)), 13, 55)
    sampler = sampling_tool_dict["vaetf"](model, dict(top_k=None, latent_dim=128, max_strlen=100, use_cond2dec=False,
                                                      decode_algo="multinomial", n_jobs=1, toklen_data=toklen_data, cond_dim=0,
                                                      scaler=None, device=dev, SRC=SRC, TRG=TRG, z_on_device=a.z_on_device))
    sampler.sample_smiles(min(a.n, 4096))                    # warm-up (workspaces, descriptors)
    t0 = time.perf_counter()
    smiles, toklen, toklen_gen = sampler.sample_smiles(a.n)
    dt = time.perf_counter() - t0
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    pd.DataFrame({"smiles": smiles, "toklen": np.asarray(toklen), "toklen_gen": toklen_gen}).to_csv(a.out)
    print(f"{a.n} SMILES in {dt:.2f} s = {a.n / dt:.0f} SMILES/s (latents drawn on the {"device" if a.z_on_device else "host, as in the reference"}); wrote {a.out}")
    print("first:", smiles[0][:60])


if __name__ == "__main__":
    main()
