"""Records batches built by the reference's OWN Model/collate_fn.py (run in the build container; /root/reference is absent
on the GPU box) into tests/golden/collate.pt.  torchtext is not installed, so SRC / TRG are oracle.collate_oracle.Field
objects (duck-typed like the fake Field of SURVEY 8c); the collate functions and their row composition are the reference's.

    python oracle/make_collate_golden.py
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GCT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import collate_oracle as CO  # noqa: E402

ATOMS = ["C", "c", "N", "n", "O", "o", "S", "s", "F", "Cl", "Br", "(", ")", "[nH]", "=", "#", "1", "2", "3", "-", "[C@@H]"]


def synthetic_frame(n, seed, props=("logP", "tPSA", "QED")):
    rng = np.random.RandomState(seed)

    def smi(lo, hi):
        return "".join(rng.choice(ATOMS, size=rng.randint(lo, hi)))

    rows = {"src": [smi(3, 40) for _ in range(n)], "src_scaffold": [smi(0, 18) for _ in range(n)]}
    rows["src"][3] = "C"                       # shortest row
    rows["src_scaffold"][5] = ""               # empty scaffold
    for p in props:
        rows[f"src_{p}"] = rng.randn(n).astype(np.float32)
        rows[f"trg_{p}"] = rng.randn(n).astype(np.float32)
    return pd.DataFrame(rows)


def main():
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_collate_fn", os.path.join(REF, "Model", "collate_fn.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    props = ["logP", "tPSA", "QED"]
    df = synthetic_frame(37, seed=3)
    # the tokeniser regex exactly as the reference's source spells it (Utils/field.py:16; the module itself imports torchtext)
    import ast
    import re
    import warnings
    line = [l for l in open(os.path.join(REF, "Utils", "field.py")) if "generaral_pattern" in l and "=" in l][0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pattern = ast.literal_eval(re.search(r'=\s*(".*")\s*$', line.strip()).group(1))
    out = {"frame": df.to_dict(orient="list"), "atoms": ATOMS, "props": props, "cases": {}, "tokenizer_pattern": pattern}
    for model_type, add_sep, use_sca, plist in (("vaetf", False, False, []), ("pvaetf", False, False, props),
                                                ("scavaetf", True, True, []), ("pscavaetf", True, True, props)):
        SRC, TRG = CO.smiles_fields(ATOMS, add_sep)
        g = torch.Generator().manual_seed(11)
        order = torch.randperm(len(df), generator=g).tolist()
        fn = ref.get_collate_fn(model_type, SRC, TRG, "cpu")
        got = []
        for i in range(0, len(order), 8):
            ins = [CO.getitem(df.iloc[r], SRC, TRG, plist, use_sca) for r in order[i:i + 8]]
            b = fn(ins)
            got.append({k: v.clone() for k, v in b.items()})
        out["cases"][model_type] = {"order": order, "batch_size": 8, "batches": got, "add_sep": add_sep, "use_scaffold": use_sca,
                                    "property_list": plist}
    path = os.path.join(ROOT, "tests", "golden", "collate.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
