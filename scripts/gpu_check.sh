#!/bin/bash
# GPU-box check used during development (run through gpurun from the repo root):
#   bash scripts/gpu_check.sh            -> pytest -m gpu, then the secondary sweep (profiles/bench_extra.py)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/bench_extra.py > gpurun_out/bench_extra.jsonl 2> gpurun_out/bench_extra.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_extra.jsonl'):
    d = json.loads(l)
    if d['kind'] == 'ntt':
        print('n=%d L=%d fwd %.3f ms (%.0f GB/s, %.2f)  inv %.3f ms (%.0f GB/s, %.2f)  pairs/s %.1fM' % (
            d['n'], d['nlimbs'], d['fwd_ms'], d['fwd_GBps'], d['fwd_frac_of_measured_hbm'], d['inv_ms'], d['inv_GBps'],
            d['inv_frac_of_measured_hbm'], d['pairs_per_s'] / 1e6))
    else:
        print('polymul n=%d B=%d %.3f ms  %.1fM products/s  %.0f GB/s' % (d['n'], d['batch'], d['ms'], d['products_per_s'] / 1e6, d['GBps']))
PY
tail -2 gpurun_out/bench_extra.err
