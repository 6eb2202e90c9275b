#!/bin/bash
# GPU-box check used during development (run through gpurun from the repo root):
#   bash scripts/gpu_check.sh [bench args]   -> pytest -m gpu, then one bench.py line (stdout -> gpurun_out/bench.json)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
except Exception as e:
    print('no bench line:', e); raise SystemExit
k = d['kernels']; r = d['roofline']
print('value %.2f M pairs/s  fwd %.4f ms  inv %.4f ms  frac %.3f  bound %s  int %.2f/%.2f' % (
    d['value'] / 1e6, k['ntt_fwd_ms'], k['ntt_inv_ms'], r['frac'], r['bound'], r['integer']['achieved'], r['integer']['peak']))
print('sustained', json.dumps(d.get('sustained')))
print('e2e %.3f M pairs/s' % (d['e2e']['value'] / 1e6), 'clocks', d['clocks'])
for kk, c in d.get('configs', {}).items():
    print(kk, '%.2f M %s' % (c['value'] / 1e6, c['unit']), 'frac %.3f' % c['roofline']['frac'], c['roofline']['bound'],
          {x: c[x] for x in ('fwd_ms', 'inv_ms', 'ms', 'launches_per_call', 'round_trip', 'vs_oracle_and_schoolbook') if x in c})
for kk, c in d.get('strong', {}).items():
    print('strong', kk, '%.2f M pairs/s' % (c['value'] / 1e6), c['round_trip'])
print('parity', d['parity_in_bench'])
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('cores'))
PY
