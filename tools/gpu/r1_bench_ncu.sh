mkdir -p gpurun_out
set -x
nproc; lscpu | grep "Model name"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -o gpurun_out/k1_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_g3 -s 6 -c 2 -o gpurun_out/k3_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
