mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "1024 2 1"; do set -- $cfg; echo "== threads $1 ctas $2 U $3"; GMG_K1_THREADS=$1 GMG_K1_CTAS=$2 GMG_K1_U=$3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'k1_ms', d['roofline']['kernel_ms'], 'k3', d['roofline']['k3_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_g3_write -s 3 -c 1 -o gpurun_out/k3w_prof_j python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
