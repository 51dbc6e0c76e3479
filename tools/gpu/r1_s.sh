mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6
timeout 300 python bench.py --workload reads400 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reads400_b.json 2> gpurun_out/bench_s.err; tail -3 gpurun_out/bench_s.err; cut -c1-2200 gpurun_out/bench_reads400_b.json
