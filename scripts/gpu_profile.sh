#!/bin/bash
# The round's evidence run on ONE B200 (through gpurun from the repo root): every command first runs plain and must exit 0,
# then under ncu.  Numbers printed by the ncu passes are never bench values.
#   bash scripts/gpu_profile.sh            -> gpurun_out/{pytest_gpu.txt, bench.json, bench_reference.json, r02_launches.csv,
#                                             prof_r02_final.ncu-rep, prof_r02_polymul.ncu-rep, prof_r02_u64.ncu-rep}
# Afterwards, here:  python profiles/summarize_ncu.py gpurun_out/prof_r02_final.ncu-rep > profiles/r02_final_ncu_summary.md  (etc.)
#                    python profiles/launch_list.py gpurun_out/r02_launches.csv > profiles/r02_final_launches.md
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err || { echo "bench FAILED"; tail -5 gpurun_out/bench.err; exit 1; }
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
LIGHT="--steps 2 --warmup 3 --no-cpu --no-extras --sustain-s 0 --e2e-steps 1"
timeout 300 python bench.py $LIGHT > gpurun_out/plain_bench.log 2>&1 || { echo "light bench FAILED"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py $LIGHT > gpurun_out/ncu_launch.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 python profiles/prof_run.py > gpurun_out/plain.log 2>&1 && \
    timeout 900 $NCU -k regex:'ntt_' -s 4 -c 2 -o gpurun_out/prof_r02_final python profiles/prof_run.py > gpurun_out/ncu.log 2>&1
timeout 300 python profiles/prof_run.py --n 2048 --polymul > gpurun_out/plain_pm.log 2>&1 && \
    timeout 900 $NCU -k regex:'ntt_' -s 3 -c 3 -o gpurun_out/prof_r02_polymul python profiles/prof_run.py --n 2048 --polymul > gpurun_out/ncu_pm.log 2>&1
timeout 300 python profiles/prof_u64.py > gpurun_out/plain_u64.log 2>&1 && \
    timeout 900 $NCU -k regex:'ref_u64' -s 1 -c 1 -o gpurun_out/prof_r02_u64 python profiles/prof_u64.py > gpurun_out/ncu_u64.log 2>&1
for nn in 1024 2048; do
  timeout 300 python profiles/prof_run.py --n $nn > gpurun_out/plain_n$nn.log 2>&1 && \
    timeout 900 $NCU -k regex:'ntt_' -s 4 -c 2 -o gpurun_out/prof_r02_n$nn python profiles/prof_run.py --n $nn > gpurun_out/ncu_n$nn.log 2>&1
done
tail -n 2 gpurun_out/plain.log gpurun_out/plain_pm.log gpurun_out/plain_u64.log gpurun_out/pytest_gpu.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
k = d['kernels']; r = d['roofline']
print('value %.2f M pairs/s  fwd %.4f ms  inv %.4f ms  frac %.3f  bound %s  int %.2f/%.2f  e2e %.3f M' % (
    d['value'] / 1e6, k['ntt_fwd_ms'], k['ntt_inv_ms'], r['frac'], r['bound'], r['integer']['achieved'], r['integer']['peak'], d['e2e']['value'] / 1e6))
for kk, c in d.get('configs', {}).items():
    print(kk, '%.2f M %s' % (c['value'] / 1e6, c['unit']), 'frac %.3f' % c['roofline']['frac'], c['roofline']['bound'])
r = json.load(open('gpurun_out/bench_reference.json'))
print('reference arm %.3f M pairs/s on %s cores' % (r['value'] / 1e6, r['cpu_baseline']['cores']))
PY
