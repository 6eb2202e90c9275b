"""Summarise an ncu report (--set full) into the handful of numbers DESIGN.md and bench.py's roofline cite.
Usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>_summary.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy (IMAD) pipe busy %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe busy %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]
print(f"# ncu summary of `{rep}`\n")
print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200).\n")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"## {d.get('Kernel Name', '?')}\n")
    print("| metric | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            print(f"| {label} (`{k}`) | {d[k]} {u.get(k, '')} |")
    rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
    stalls = sorted(((float(v), k) for k, v in d.items()
                     if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v),
                    reverse=True)
    print("\nTop warp stall reasons (warps stalled per issued instruction):\n")
    for v, k in stalls[:8]:
        print(f"- {k.split('stalled_')[1].split('_per_issue')[0]}: {v:.3f}")
    print()
