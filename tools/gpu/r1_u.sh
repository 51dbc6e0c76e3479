mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 300 python tools/gpu/train_levels.py 1.0 2>&1 | tail -10
timeout 300 python bench.py --workload train500m --steps 10 --warmup 3 > gpurun_out/bench_train500m.json 2> gpurun_out/bench_u.err; tail -3 gpurun_out/bench_u.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_train500m.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'], d['config']['model_sha256_16'])
PY
