mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -3 gpurun_out/bench_r.err; cat gpurun_out/bench_r.json | cut -c1-1500
for i in 1 2 3; do CUDA_LAUNCH_BLOCKING=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/blk$i.log 2>&1; tail -1 gpurun_out/blk$i.log | cut -c1-120; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -f -o gpurun_out/k1_prof_r python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log | cut -c1-200
