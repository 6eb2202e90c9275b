run() { python bench.py --steps 30 --warmup 5 --no-cpu --e2e-steps 0 > gpurun_out/t.json 2>gpurun_out/t.err; python -c "
import json; d=json.load(open('gpurun_out/t.json')); print('$1 fwd_ms %.4f inv_ms %.4f' % (d['kernels']['ntt_fwd_ms'], d['kernels']['ntt_inv_ms']), d['parity_in_bench'])"; }
run default
AGX_CARVEOUT_FWD=77 AGX_CARVEOUT_INV=60 run fwd77_inv60
AGX_CARVEOUT_FWD=64 AGX_CARVEOUT_INV=52 run fwd64_inv52
AGX_CARVEOUT_FWD=100 AGX_CARVEOUT_INV=100 run fwd100_inv100
AGX_LIB=$PWD/agilex-ntt_b200/lib/libagxntt_old.so run old
