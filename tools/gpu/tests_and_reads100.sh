# the GPU test suite, then reads100 with and without the codon-per-lane kernel (A/B)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=3 ) > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2x_tests.log
( timeout 600 python bench.py --workload reads100 --steps 32 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2x_reads100.json 2> gpurun_out/r2x_reads100.err; echo "reads100 rc=$?"; tail -c 300 gpurun_out/r2x_reads100.err
( GMG_PLAIN_LANES=0 timeout 600 python bench.py --workload reads100 --steps 32 --warmup 5 --no-cpu-baseline --no-extra --no-pipeline ) > gpurun_out/r2x_reads100_nolanes.json 2> gpurun_out/r2x_reads100_nolanes.err; echo "nolanes rc=$?"
python - <<'PY'
import json
for f in ('r2x_reads100','r2x_reads100_nolanes'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),x['e2e'].get('ms_per_step'),'one',x['e2e'].get('one_batch_at_a_time'),'k',x['roofline'].get('ms_per_step_by_kernel'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
