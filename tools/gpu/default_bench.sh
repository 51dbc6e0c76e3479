# the default bench (N=1, everything on) and the reference arm, as the driver runs them
mkdir -p gpurun_out
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2y_bench_ref.json 2> gpurun_out/r2y_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2y_bench_ref.json
( timeout 1200 python bench.py ) > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; echo "bench n1 rc=$?"; tail -c 600 gpurun_out/r2y_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_n1.json') if l.startswith('{')][-1])
def show(tag,x):
    print(tag,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),round(x['e2e'].get('ms_per_step',0),3),'roof',x['roofline'].get('kernel'),round(x['roofline']['frac'],3),x['roofline'].get('traffic'),'parity',x.get('parity_checked'),'cpu',x.get('cpu_baseline',{}).get('value'),'app',(x.get('e2e_app') or {}).get('speedup'))
show('reads100',d)
for k,v in d.get('extra',{}).items(): show(k,v)
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
