mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 500 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5
timeout 300 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -3 gpurun_out/bench_m.err; cut -c1-2500 gpurun_out/bench_m.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_m_ref.json 2> gpurun_out/bench_m_ref.err; tail -3 gpurun_out/bench_m_ref.err; cut -c1-1500 gpurun_out/bench_m_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -f -o gpurun_out/k1_prof_m python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
