mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; cut -c1-1500 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_final_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -f -o gpurun_out/k1_prof_final python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_mg_starts_warp -s 2 -c 2 -f -o gpurun_out/k3mg_prof python bench.py --workload reads400 --steps 1 --warmup 3 --scale 0.2 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_prefix_lanes -s 2 -c 1 -f -o gpurun_out/k2_prof python bench.py --workload reads100 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
for w in reads400 reads100 train500m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -2 gpurun_out/bench_$w.err; cut -c1-200 gpurun_out/bench_$w.json
done
