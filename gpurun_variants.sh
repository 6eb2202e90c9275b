mkdir -p gpurun_out
nvidia-smi --query-gpu=uuid,serial --format=csv,noheader
for rep in 1 2; do
for v in "" _swz; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  python bench.py --steps 50 --warmup 5 --no-cpu --e2e-steps 1 > gpurun_out/bench$v.json 2>gpurun_out/bench$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench$v.json')); print('variant[$v] fwd_ms %.4f inv_ms %.4f' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms']), d['clocks'])"
done
done
