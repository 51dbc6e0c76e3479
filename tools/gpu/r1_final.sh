mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_final.json')); print('contig5m', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
PY
for w in reads400 reads100 train500m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -2 gpurun_out/bench_$w.err; python - <<PY
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value'],3), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],3), round(d['e2e']['ms_per_step'],3), d['roofline'].get('ms_per_step_by_kernel'), round(d['roofline']['frac'],3))
PY
done
