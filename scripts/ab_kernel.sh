#!/bin/bash
# A/B of the two kernel families inside ONE gpurun call (the pool's B200s differ by ~5 %):  bash scripts/ab_kernel.sh [variants...]
# a variant is a value of AGX_KERNEL (2p, r16); every run is wrapped in a timeout
mkdir -p gpurun_out
vars=${@:-2p r16}
for v in $vars; do
  AGX_KERNEL=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "4096 or config2 or config3 or caller or survey or golden or edge or reduced or polymul or multi or shards" 2>&1 | tail -3
done
for rep in 1 2; do
for v in $vars; do
  AGX_KERNEL=$v timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 --no-extras --sustain-s 1.0 > gpurun_out/bench_ab_$v.json 2>gpurun_out/bench_ab_$v.err || { echo "variant[$v] FAILED or timed out"; tail -3 gpurun_out/bench_ab_$v.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$v.json')); s=d['sustained']; print('variant[$v] %s fwd_ms %.4f inv_ms %.4f | sustained %.2f M pairs/s @ %s MHz %s W' % (d['kernel_variant'], d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms'], s['value']/1e6, s['clocks']['sm_mhz'], s['clocks']['power_w_median']), d['parity_in_bench']['forward_vs_oracle'], d['parity_in_bench']['round_trips_and_e2e'])"
done
done
