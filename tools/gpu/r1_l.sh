mkdir -p gpurun_out
./tools/gpu/ubench_leaf 2>&1 | tee gpurun_out/ubench_leaf.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -3 gpurun_out/bench_l.err; cat gpurun_out/bench_l.json
