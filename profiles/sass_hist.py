"""SASS opcode histogram of every kernel in libagxntt.so (static counts; the looped kernels execute their stage code
twice per transform).  Usage (no GPU needed): python profiles/sass_hist.py > profiles/rNN_sass_histograms.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "agilex-ntt_b200", "lib", "libagxntt.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print(f"# SASS opcode histograms of `{os.path.relpath(lib, ROOT)}` (sm_100a)\n")
print("`cuobjdump -sass`, static instruction counts per kernel; opcodes with their first modifier.  `UTMASTG` / `UTMALDG` are the")
print("TMA tensor store / load, `IMAD.HI` is the half-rate multiply that sets the integer roofline (one per butterfly).\n")
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("agx::", "")
    ops = collections.Counter()
    n = 0
    for l in part.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)?)", l)
        if m:
            op = m.group(1)
            if op.startswith(("IMAD", "UTMA", "LDG", "STG", "LDS", "STS", "BAR")):
                pass
            else:
                op = op.split(".")[0]
            ops[op] += 1
            n += 1
    top = ", ".join(f"{k} {v}" for k, v in ops.most_common(14))
    tma = {k: v for k, v in ops.items() if k.startswith("UTMA") or k.startswith("UBLKCP")}
    print(f"* `{dem}` -- {n} instructions ({n * 16 / 1024:.1f} KB): {top}" + (f"; **TMA: {tma}**" if tma else ""))
