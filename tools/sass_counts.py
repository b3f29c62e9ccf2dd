"""SASS evidence per tensor-core kernel of the shipped library -> profiles/r02_sass.md
(python tools/sass_counts.py; needs cuobjdump and c++filt, no GPU)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "keras_unsupervised_b200", "libkucd.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = {"name": m.group(1), "lines": 0, "mma": 0, "mma2": 0, "tma": 0, "ldtm": 0, "bar": 0, "hmma": 0}
        kernels.append(cur)
        continue
    if cur is None or "/*" not in line:
        continue
    cur["lines"] += 1
    if "UTCHMMA" in line:
        cur["mma"] += 1
        cur["mma2"] += ".2CTA" in line
    elif re.search(r"\bHMMA\b", line):
        cur["hmma"] += 1
    if "UTMALDG" in line:
        cur["tma"] += 1
    if "LDTM" in line:
        cur["ldtm"] += 1
    if "UTCBAR" in line:
        cur["bar"] += 1
names = subprocess.run(["c++filt"], input="\n".join(k["name"] for k in kernels), capture_output=True, text=True).stdout.splitlines()
for k, n in zip(kernels, names):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*\)$", "", n).replace("kucd::", "").replace("(bool)1", "true").replace("(bool)0", "false")
    k["pretty"] = re.sub(r"\(int\)", "", n)
tc = [k for k in kernels if k["mma"]]
tot = {f: sum(k[f] for k in kernels) for f in ("mma", "mma2", "tma", "ldtm", "bar", "hmma")}
out = ["# SASS evidence per tensor-core kernel (round 2, final build)", "",
       "`cuobjdump -sass keras_unsupervised_b200/libkucd.so` of the shipped build (`nvcc -gencode arch=compute_100a,code=sm_100a",
       "-lineinfo -O3`), counted per kernel by `tools/sass_counts.py`: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG` = TMA",
       "tensor loads (cp.async.bulk.tensor), `LDTM` = tcgen05.ld (tensor memory -> registers), `UTCBAR` = tcgen05.commit",
       "(mbarrier arrive, multicast for CTA pairs).  %d kernels in the library, %d of them tensor-core kernels; totals: %d" % (len(kernels), len(tc), tot["mma"]),
       "`UTCHMMA` (%d `.2CTA`), %d `UTMALDG`, %d `LDTM`, %d `UTCBAR`, **%d legacy `HMMA`** (no mma.sync / wmma anywhere)." % (tot["mma2"], tot["tma"], tot["ldtm"], tot["bar"], tot["hmma"]),
       "The probe hooks (`dbg_flags`, `dbg_lbo_*`) are compiled only into `build/probe_gemm` (`#ifdef KUCD_PROBE`): the product",
       "kernels have no parameter that can skip a load or an MMA.  Template arguments: `chain_kernel<BN, CG, GAUSS, CH>`,",
       "`gemm_bf16_kernel<BN, A_MN, B_MN, EPI, CH, CG>` (CH = 8: float32-grade piecewise accumulation).", "",
       "| kernel | SASS lines | UTCHMMA (of which .2CTA) | UTMALDG | LDTM | UTCBAR | legacy HMMA |", "|---|---|---|---|---|---|---|"]
for k in tc:
    out.append("| `%s` | %d | %d (%d) | %d | %d | %d | %d |" % (k["pretty"], k["lines"], k["mma"], k["mma2"], k["tma"], k["ldtm"], k["bar"], k["hmma"]))
open(os.path.join(ROOT, "profiles", "r02_sass.md"), "w").write("\n".join(out) + "\n")
print("%d kernels, %d tensor-core kernels, totals %s" % (len(kernels), len(tc), tot), file=sys.stderr)
