timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
for w in reads100 reads400; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err; tail -2 gpurun_out/bench_t2.err; python - <<PY
import json; d=json.load(open('gpurun_out/bench_t2.json')); print('$w', round(d['value'],3), round(d['ms_per_step'],3), d['roofline']['ms_per_step_by_kernel'], round(d['roofline']['frac'],3))
PY
done
