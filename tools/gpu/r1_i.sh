mkdir -p gpurun_out
nproc; lscpu | grep "Model name"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -3 gpurun_out/bench_i.err; cat gpurun_out/bench_i.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_i.json 2>&1; cat gpurun_out/bench_ref_i.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -o gpurun_out/k1_prof_i python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out
