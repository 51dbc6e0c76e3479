mkdir -p gpurun_out
for i in 1 2; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_planes -s 3 -c 1 -f -o gpurun_out/k1_prof_q python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bq$i.log 2>&1; grep -v "^$" gpurun_out/ncu_bq$i.log | tail -2 | cut -c1-300
done
for i in 1 2 3 4 5 6; do
CUDA_LAUNCH_BLOCKING=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/blk$i.log 2>&1; tail -1 gpurun_out/blk$i.log | cut -c1-200
done
