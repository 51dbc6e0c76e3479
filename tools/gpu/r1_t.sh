timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; tail -2 gpurun_out/bench_t.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_t.json')); print('contig5m', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['k3_ms_per_step'])
PY
done
