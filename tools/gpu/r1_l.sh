mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -3 gpurun_out/bench_p.err; cat gpurun_out/bench_p.json
GMG_K1_U=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('U1', d['value'], d['ms_per_step'], 'k1_ms', d['roofline']['kernel_ms'], 'k3', d['roofline']['k3_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -o gpurun_out/k1_prof_p python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
