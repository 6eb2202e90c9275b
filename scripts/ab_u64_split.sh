#!/bin/bash
# N = 16384 device-resident: one CTA per frame (AGX_REF_SPLIT14=0) against two (=1) and the adaptive default, one call
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ref_pipeline.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for mode in 0 1 default; do
  if [ $mode = default ]; then unset AGX_REF_SPLIT14; else export AGX_REF_SPLIT14=$mode; fi
  timeout 300 python profiles/bench_u64_dev.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d.get('N') == 16384: print('split14=$mode', d['chunk_MiB'], 'MiB', round(d['frames_per_s'] / 1e6, 3), 'M frames/s', d['parity'])"
done
done
