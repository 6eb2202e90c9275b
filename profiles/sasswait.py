"""List a kernel's SASS with the scoreboard fields of every instruction: which scoreboard a load SETS (write barrier) and which
scoreboards an instruction WAITS for.  Shows where the first consumer of a batch of loads stalls.
Usage: cuobjdump -sass lib.so | python profiles/sasswait.py <kernel-name-substring> [start_hex end_hex]"""
import re
import sys

txt = sys.stdin.read()
want = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if want not in name:
        continue
    lines = part.split("\n")
    for i, l in enumerate(lines):
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", l)
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if not m2:
                continue
            addr = int(m.group(1), 16)
            if not (lo <= addr < hi):
                continue
            hiw = int(m2.group(1), 16)
            stall = (hiw >> 41) & 0xF
            wbar = (hiw >> 46) & 0x7      # scoreboard set when the result is written (7 = none)
            rbar = (hiw >> 49) & 0x7      # scoreboard set when the operands have been read (7 = none)
            wmask = (hiw >> 52) & 0x3F    # scoreboards waited for before issue
            waits = ",".join(str(b) for b in range(6) if wmask >> b & 1)
            print(f"{addr:05x} st{stall:2d} set{'-' if wbar == 7 else wbar} rd{'-' if rbar == 7 else rbar} wait[{waits:6s}] {m.group(2)[:70]}")
    break
