"""Decode the control words of a kernel's SASS: per-instruction stall count, yield, scoreboard waits.
Prints the sum of stall counts (= minimum cycles for ONE warp to issue the code in isolation) per opcode class.
Usage: cuobjdump -sass lib.so | python profiles/sassctl.py <kernel-name-substring> [start_hex end_hex]"""
import re
import sys
import collections

txt = sys.stdin.read()
want = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if want not in name:
        continue
    lines = part.split("\n")
    ins = []
    for i, l in enumerate(lines):
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+).*?/\* 0x([0-9a-f]{16}) \*/", l)
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                addr = int(m.group(1), 16)
                hiw = int(m2.group(1), 16)
                stall = (hiw >> 41) & 0xF
                yld = (hiw >> 45) & 1
                wmask = (hiw >> 52) & 0x3F
                if lo <= addr < hi:
                    ins.append((addr, m.group(2), stall, yld, wmask))
    tot = sum(s for _, _, s, _, _ in ins)
    print(name[:70], "instr", len(ins), "sum(stall)", tot, "avg %.2f" % (tot / max(len(ins), 1)))
    by = collections.defaultdict(lambda: [0, 0])
    for _, op, s, _, _ in ins:
        k = op if op.startswith("IMAD") else op.split(".")[0]
        by[k][0] += 1
        by[k][1] += s
    for k, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k:16s} n={n:5d} stall_sum={s:6d} avg={s / n:.2f}")
    hist = collections.Counter(s for _, _, s, _, _ in ins)
    print("   stall histogram:", dict(sorted(hist.items())))
