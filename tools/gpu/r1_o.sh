mkdir -p gpurun_out
timeout 300 python tools/gpu/train_levels.py 1.0 2>&1 | tail -12
for w in reads400 reads100; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$w.csv | tail -25
done
