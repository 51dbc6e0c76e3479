# round 2, GPU call R: bucket_finish from three words, offsets copied from the caller's buffer, 2 / 3 / 4 batches in flight
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=3 ) > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2r_tests.log
for L in 2 3 4; do
( GMG_BENCH_LANES=$L timeout 600 python bench.py --workload reads100 --steps 24 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2r_reads100_l$L.json 2> gpurun_out/r2r_reads100_l$L.err; echo "lanes $L rc=$?"; tail -c 300 gpurun_out/r2r_reads100_l$L.err
done
( GMG_BENCH_LANES=3 timeout 600 python bench.py --workload reads400 --steps 12 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2r_reads400.json 2> gpurun_out/r2r_reads400.err; echo "reads400 rc=$?"
python tools/gpu/e2e_breakdown_reads.py 2>&1 | tail -11
python - <<'PY'
import json
for f in ('r2r_reads100_l2','r2r_reads100_l3','r2r_reads100_l4','r2r_reads400'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),x['e2e'].get('ms_per_step'),'one',x['e2e'].get('one_batch_at_a_time'),'k',x['roofline'].get('ms_per_step_by_kernel'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
