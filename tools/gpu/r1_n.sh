mkdir -p gpurun_out
for w in reads400 reads100 train500m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -5 gpurun_out/bench_$w.err; cut -c1-3000 gpurun_out/bench_$w.json
  timeout 300 python bench.py --impl reference --workload $w --steps 2 --warmup 1 > gpurun_out/bench_${w}_ref.json 2> gpurun_out/bench_${w}_ref.err
  tail -3 gpurun_out/bench_${w}_ref.err; cut -c1-600 gpurun_out/bench_${w}_ref.json
done
