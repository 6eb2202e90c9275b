mkdir -p gpurun_out
for v in _ds; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fwd_inv_vs_oracle" 2>&1 | tail -1
  python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench$v.json 2>gpurun_out/bench$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench$v.json')); print('variant[$v] fwd_ms', d['kernels']['ntt_fwd_ms'], 'inv_ms', d['kernels']['ntt_inv_ms'], d['parity_in_bench'])"
done
