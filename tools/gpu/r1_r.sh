mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -3 gpurun_out/bench_r.err; cut -c1-1800 gpurun_out/bench_r.json
python tools/gpu/e2e_breakdown.py 2>&1 | tail -10
