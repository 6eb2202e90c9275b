mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt_trace.so python profiles/trace_phases.py 2>&1 | tee gpurun_out/trace_phases.txt | head -3
for v in "" _pf2; do
  export AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt$v.so
  python bench.py --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench$v.json 2>gpurun_out/bench$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench$v.json')); print('variant[$v] fwd_ms', d['kernels']['ntt_fwd_ms'], 'inv_ms', d['kernels']['ntt_inv_ms'])"
done
