# round 2, GPU call C: whole GPU suite on the flat K3 / reduction / chunk-level binding, then reads400 + default bench
mkdir -p gpurun_out
( timeout 2000 python -m pytest tests -m gpu -q -x --timeout 900 --durations=8 ) > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/r2c_tests.log
( timeout 600 python bench.py --workload reads400 --steps 10 --warmup 3 ) > gpurun_out/r2c_reads400.json 2> gpurun_out/r2c_reads400.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c_reads400.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_reads400.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'kernels',d['roofline'].get('ms_per_step_by_kernel'),'\nparity',d['parity'],'\napp',d.get('e2e_app'))
except Exception as e: print('no json', e)
PY
