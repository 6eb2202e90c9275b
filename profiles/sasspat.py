"""Print the pipe-class pattern of a kernel's SASS (M = IMAD-pipe op, H = IMAD.HI/WIDE, a = ALU op, l = memory,
. = other) and the run-length statistics of M/H runs.  Usage: cuobjdump -sass x | python profiles/sasspat.py [name-substr]"""
import re
import sys
import collections

txt = sys.stdin.read()
want = sys.argv[1] if len(sys.argv) > 1 else ""
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if want not in name:
        continue
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", part)
    pat = []
    for o in ops:
        b = o.split(".")[0]
        if o.startswith("IMAD.HI") or o.startswith("IMAD.WIDE"):
            pat.append("H")
        elif b == "IMAD":
            pat.append("M")
        elif b in ("IADD3", "VIADDMNMX", "LOP3", "LEA", "SHF", "VIADD", "SEL", "ISETP", "MOV", "PRMT", "IABS"):
            pat.append("a")
        elif b in ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "LD", "ST"):
            pat.append("l")
        else:
            pat.append(".")
    s = "".join(pat)
    runs = [len(m.group(0)) for m in re.finditer(r"[MH]+", s)]
    hist = collections.Counter(min(r, 33) for r in runs)
    print(name[:70], len(ops), "instr; M/H runs:", len(runs), "mean %.1f" % (sum(runs) / max(len(runs), 1)), "max", max(runs or [0]))
    if "-v" in sys.argv:
        for i in range(0, len(s), 128):
            print(s[i:i + 128])
