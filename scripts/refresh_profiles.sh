#!/bin/bash
# After `gpurun -- bash scripts/gpu_profile.sh`: turn what came back in gpurun_out/ into the tracked files under profiles/.
set -e
cd "$(dirname "$0")/.."
python profiles/summarize_ncu.py gpurun_out/prof_r02_final.ncu-rep > profiles/r02_final_ncu_summary.md
python profiles/summarize_ncu.py gpurun_out/prof_r02_polymul.ncu-rep > profiles/r02_polymul_n2048_ncu_summary.md
python profiles/summarize_ncu.py gpurun_out/prof_r02_u64.ncu-rep > profiles/r02_u64_frame_ncu_summary.md
for nn in 1024 2048; do
  [ -f gpurun_out/prof_r02_n$nn.ncu-rep ] && python profiles/summarize_ncu.py gpurun_out/prof_r02_n$nn.ncu-rep > profiles/r02_n${nn}_ncu_summary.md
done
cp gpurun_out/r02_launches.csv profiles/r02_final_launches.csv
cp gpurun_out/bench.json profiles/r02_c_bench_1gpu.json
cp gpurun_out/bench_reference.json profiles/r02_c_bench_reference.json
cp gpurun_out/pytest_gpu.txt profiles/r02_c_pytest_gpu.txt
python profiles/sass_hist.py > profiles/r02_sass_histograms.md
python - <<'PY'
import json, re, subprocess
tab = subprocess.run(["python", "profiles/launch_list.py", "profiles/r02_final_launches.csv"], capture_output=True, text=True).stdout
rows = [[x.strip() for x in l.strip("|").split("|")] for l in tab.splitlines() if l.startswith("| `")]
mean = lambda name: next(float(c[4]) for c in rows if name in c[0] and c[1] == "65536")
f, i = mean("ntt_fwd_loop_kernel<12"), mean("ntt_inv_loop_kernel<12")
b = json.load(open("profiles/r02_c_bench_1gpu.json"))
kf, ki = b["kernels"]["ntt_fwd_ms"], b["kernels"]["ntt_inv_ms"]
open("profiles/r02_final_launches.md", "w").write(f"""# Launch list of one `bench.py` run, final round-2 build (`profiles/r02_final_launches.csv`)

`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` over
`python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --sustain-s 0 --e2e-steps 1` (the same command exited 0 without
ncu first; `scripts/gpu_profile.sh`).  Times are cold-cache and serialised by the profiler: compare SHARES, not absolutes.
Every kernel is the repo's own (two `at::` helpers come from the torch `==` of bench.py's end-to-end spot check).

{tab}
The headline step = one forward + one inverse launch with grid 65 536 (the whole 1 GiB batch): forward {f:.1f} us
({100 * f / (f + i):.1f} % of the step), inverse {i:.1f} us ({100 * i / (f + i):.1f} %) -- the CUDA-event timings inside
bench.py (`profiles/r02_c_bench_1gpu.json`: {kf:.4f} / {ki:.4f} ms) split {100 * kf / (kf + ki):.1f} / {100 * ki / (kf + ki):.1f}.  The grid-2048 launches are the 32 MiB
chunks of the end-to-end host pipeline (`agx_ntt_fwd_host` / `agx_ntt_inv_host`); `diag_bfly_kernel` is
`agx_measure_butterfly_peak` (the measured integer roofline), `checksum_kernel` / `fill_synthetic_kernel` are the parity
checks and the synthetic input.
""")
txt = open("profiles/r02_final_ncu_summary.md").read()
def dram(kern):
    sec = txt[txt.index(kern):]
    g = lambda pat: float(re.search(pat + r".*?\| ([0-9.]+) Gbyte", sec).group(1))
    return round((g("DRAM bytes read") + g("DRAM bytes written")) * 1e9)
json.dump({"_source": "profiles/r02_final_ncu_summary.md (ncu --set full --clock-control none, per launch of 65,536 n=4096 transforms, final round-2 build; dram__bytes_read.sum + dram__bytes_write.sum)",
           "ntt_fwd_kernel": dram("ntt_fwd_loop_kernel"), "ntt_inv_kernel": dram("ntt_inv_loop_kernel")},
          open("profiles/roofline_traffic.json", "w"), indent=1)
print(open("profiles/roofline_traffic.json").read())
PY
git status --short | head
