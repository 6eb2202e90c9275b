#!/bin/bash
# A/B of run-time knobs inside ONE gpurun call: bash scripts/ab_env.sh "" "AGX_CARVEOUT=58" ...   (each argument = env assignments)
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  env $v timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 --no-extras --sustain-s 1.0 > gpurun_out/bench_env.json 2>gpurun_out/bench_env.err || { echo "[$v] FAILED or timed out"; tail -2 gpurun_out/bench_env.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/bench_env.json')); print('[$v] fwd_ms %.4f inv_ms %.4f sustained %.2f M' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms'], d['sustained']['value']/1e6), d['parity_in_bench']['forward_vs_oracle'])"
done
done
