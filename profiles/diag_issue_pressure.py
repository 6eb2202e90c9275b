"""Issue-pressure / register-read experiments on the butterfly stream (agx_measure_butterfly_peak kinds 0, 2-6):
butterflies per clock per SM at 1, 2, 3, 4, 6, 8 warps per scheduler.  Usage (GPU box): python profiles/diag_issue_pressure.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agilex_ntt_b200 as A
c = A.Context(1024, [1053818881])
names = {0: "butterfly (6 instr)", 2: "+1 LOP3 reading 3 registers", 3: "+2 LOP3 reading 3 registers", 4: "+1 LOP3 reading 1 register",
         5: "twiddle from the constant bank", 6: "distinct twiddle registers per chain"}
for kind in (0, 2, 3, 4, 5, 6):
    print(json.dumps({"kind": kind, "what": names[kind], "threads_per_sm": [128, 256, 384, 512, 768, 1024],
                      "butterflies_per_clk_per_sm": [round(c.measure_butterfly_peak(kind, t)[0], 2) for t in (128, 256, 384, 512, 768, 1024)]}))
