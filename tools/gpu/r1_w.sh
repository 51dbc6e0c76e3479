mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_prefix_lanes -s 2 -c 1 -f -o gpurun_out/k2_prof python bench.py --workload reads100 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log | cut -c1-200
timeout 300 python bench.py --workload reads100 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reads100_c.json 2> gpurun_out/bench_v.err; tail -3 gpurun_out/bench_v.err; python - <<PY
import json; d=json.load(open('gpurun_out/bench_reads100_c.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
