#!/bin/bash
# A/B of u64-path builds inside ONE gpurun call: bash scripts/ab_u64.sh "" _u64old _u64negq
# (variant "_x" = agilex-ntt_b200/lib/libagxntt_x.so, "" = the shipped library); parity first, then the device-resident rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ref_pipeline.py -m gpu -x -q 2>&1 | tail -3
[ -x experiments/bin/u64_bfly_variants ] && timeout 120 experiments/bin/u64_bfly_variants | tee gpurun_out/u64_bfly_variants.jsonl
for rep in 1 2; do
for v in "$@"; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  timeout 300 python profiles/bench_u64_dev.py > gpurun_out/bench_u64_dev$v.$rep.jsonl 2> gpurun_out/bench_u64_dev$v.err || { echo "variant[$v] FAILED"; tail -3 gpurun_out/bench_u64_dev$v.err; continue; }
  echo "variant[$v] rep $rep"; python - <<PY
import json
for l in open('gpurun_out/bench_u64_dev$v.$rep.jsonl'):
    d = json.loads(l)
    if 'N' in d: print('  ', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('N', 'chunk_MiB', 'frames_per_s', 'parity', 'launches', 'frac_of_measured_u64_integer_peak', 'ms')})
    else: print('  ', d)
PY
done
done
