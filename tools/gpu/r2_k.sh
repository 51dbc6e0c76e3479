# round 2, GPU call K: full GPU suite + default bench (N=1)
mkdir -p gpurun_out
( timeout 2000 python -m pytest tests -m gpu -q -x --timeout 900 --durations=5 ) > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2k_tests.log
( timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench n1 rc=$?"; tail -c 1000 gpurun_out/r2k_bench_n1.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2k_bench_ref.json
python - <<'PY'
import json
for f in ('r2k_bench_n1',):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        def show(tag,x):
            print(f, tag,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),'e2e_ms',round(x['e2e'].get('ms_per_step',0),3),'d2h',x['e2e']['d2h_bytes_per_step'],'by_kernel',x['roofline'].get('ms_per_step_by_kernel'),x['roofline'].get('kernel'),round(x['roofline']['frac'],3),'parity',x.get('parity_checked'), 'app', (x.get('e2e_app') or {}).get('speedup'))
        show('reads100',d)
        for k,v in d.get('extra',{}).items(): show(k,v)
    except Exception as e: print(f,'no json',e)
PY
