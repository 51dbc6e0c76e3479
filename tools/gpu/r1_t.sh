timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; tail -2 gpurun_out/bench_t.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_t.json')); print('contig5m', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])
PY
for w in reads400 reads100; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err; tail -2 gpurun_out/bench_t2.err; python - <<PY
import json; d=json.load(open('gpurun_out/bench_t2.json')); print('$w', round(d['value'],3), round(d['ms_per_step'],3), d['roofline']['ms_per_step_by_kernel'])
PY
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "score_all_frames or many_contigs or mg_start_lists_match_reference" 2>&1 | grep "=========\|passed\|failed" | head -6
