mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_final.json')); print('contig5m', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; cut -c1-120 gpurun_out/bench_final_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -f -o gpurun_out/k1_prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-100
