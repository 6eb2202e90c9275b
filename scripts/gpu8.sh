#!/bin/bash
# One 8-GPU call: raw PCIe ceiling of the box, the hardware shard-and-compare test on all GPUs, bench.py under torchrun,
# and the end-to-end pipeline's sensitivity to its chunk size at 8 ranks.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" > gpurun_out/lscpu_$N.txt 2>&1
timeout 300 $TR --master-port 29511 profiles/pcie_ceiling.py > gpurun_out/pcie_ceiling_$N.json 2> gpurun_out/pcie_ceiling_$N.err; echo "pcie rc=$?"; cat gpurun_out/pcie_ceiling_$N.json
timeout 300 python profiles/pcie_ceiling.py > gpurun_out/pcie_ceiling_1of$N.json 2>> gpurun_out/pcie_ceiling_$N.err; cat gpurun_out/pcie_ceiling_1of$N.json
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_multi_gpu_$N.txt
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$N.err
for mb in 16 64 128; do
  AGX_HOST_CHUNK_MB=$mb timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-extras --sustain-s 0 --e2e-steps 5 > gpurun_out/bench_${N}_chunk$mb.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_${N}_chunk$mb.json')); print('chunk $mb MiB: e2e %.3f M pairs/s' % (d['e2e']['value']/1e6))"
done
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$N.json'))
print('N=$N value %.1f M pairs/s e2e %.3f M' % (d['value']/1e6, d['e2e']['value']/1e6), d['parity_in_bench'])
print('sustained', d['sustained']['value']/1e6, d['sustained']['clocks'])
for k, c in d['strong'].items(): print('strong', k, '%.1f M' % (c['value']/1e6), c['round_trip'])
for k, c in d['configs'].items(): print(k, '%.1f M %s' % (c['value']/1e6, c['unit']))
PY
