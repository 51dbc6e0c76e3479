# round 2, GPU call B: flat K3 + reduction on the device
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 900 -k "mg_ or config3 or config5 or smoke" --durations=8 ) > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/r2b_tests.log
( timeout 600 python bench.py --workload reads400 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2b_reads400.json 2> gpurun_out/r2b_reads400.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2b_reads400.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2b_reads400.json'))
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'kernels',d['roofline'].get('ms_per_step_by_kernel'),'parity',d['parity'])
except Exception as e: print('no json', e)
PY
