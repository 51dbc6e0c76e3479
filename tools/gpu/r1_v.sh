mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4
for w in reads100 reads400; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_c.json 2> gpurun_out/bench_v.err; tail -3 gpurun_out/bench_v.err; python - <<PY
import json; d=json.load(open('gpurun_out/bench_${w}_c.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
done
