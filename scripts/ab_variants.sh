mkdir -p gpurun_out
for v in "" _t320 _t384 _t448; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench$v.json 2>gpurun_out/bench$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench$v.json')); print('variant[$v] fwd_ms %.4f inv_ms %.4f' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms']), d['parity_in_bench'])"
done
