mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5
timeout 300 python bench.py --workload reads400 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reads400_b.json 2> gpurun_out/bench_s.err; tail -3 gpurun_out/bench_s.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_reads400_b.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kernel'], d['e2e']['ms_per_step'], d['config']['starts_last_step'])
PY
