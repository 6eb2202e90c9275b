#!/bin/bash
# A/B of experiment builds inside ONE gpurun call (the pool's B200s differ by ~5 %): bash scripts/ab_variants.sh "" _bulk ...
# every run is wrapped in a timeout so that a kernel that hangs costs a minute, not the call's whole limit
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  timeout 90 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench_ab$v.json 2>gpurun_out/bench_ab$v.err || { echo "variant[$v] FAILED or timed out"; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab$v.json')); print('variant[$v] fwd_ms %.4f inv_ms %.4f' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms']), d['parity_in_bench'])"
done
done
