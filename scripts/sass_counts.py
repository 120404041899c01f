"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (1-D bulk copy), HMMA (mma.sync), MOVM (movmatrix).
    python scripts/sass_counts.py [path/to/libgct_b200.so] > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gct_plus_b200", "libgct_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
MN = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MOVM", "SYNCS"]
counts, cur, order = {}, None, []
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it))
        if cur not in counts:
            counts[cur] = collections.Counter()
            order.append(cur)
        continue
    if cur is None:
        continue
    for k in MN:
        if k == "UTCHMMA.2CTA":
            if "UTCHMMA.2CTA" in line:
                counts[cur][k] += 1
        elif re.search(r"\b" + re.escape(k) + r"\b", line.replace(".", " ")):
            counts[cur][k] += 1
tot = collections.Counter()
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); kernels without any are omitted")
print(f"{'kernel':92s} " + " ".join(f"{k:>12s}" for k in MN))
for n in order:
    c = counts[n]
    if not any(c.values()):
        continue
    tot.update(c)
    print(f"{n[:92]:92s} " + " ".join(f"{c[k]:12d}" for k in MN))
print(f"{'TOTAL':92s} " + " ".join(f"{tot[k]:12d}" for k in MN))
