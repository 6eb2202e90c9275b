mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench.json 2>gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms'], d['roofline']['frac'], d['clocks'])"
