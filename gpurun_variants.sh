mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_full.json 2>gpurun_out/bench_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_full.json; cat gpurun_out/bench_ref.json | cut -c1-300
