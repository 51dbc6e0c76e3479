mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
./tools/gpu/ubench_gather 2>&1 | tee gpurun_out/ubench_gather.txt
python tools/gpu/e2e_breakdown.py 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -3 gpurun_out/bench_e.err; cat gpurun_out/bench_e.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
