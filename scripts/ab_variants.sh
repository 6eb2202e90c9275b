#!/bin/bash
# A/B of experiment builds inside ONE gpurun call (the pool's B200s differ by ~5 %): bash scripts/ab_variants.sh "" _r112 ...
# (a variant "_x" is agilex-ntt_b200/lib/libagxntt_x.so; "" is the shipped library); every run is wrapped in a timeout
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 --no-extras --sustain-s 1.0 > gpurun_out/bench_ab$v.json 2>gpurun_out/bench_ab$v.err || { echo "variant[$v] FAILED or timed out"; tail -2 gpurun_out/bench_ab$v.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab$v.json')); print('variant[$v] fwd_ms %.4f inv_ms %.4f sustained %.2f M' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms'], d['sustained']['value']/1e6), d['parity_in_bench']['forward_vs_oracle'], d['parity_in_bench']['round_trips_and_e2e'])"
done
done
