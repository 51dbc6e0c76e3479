mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k3_g3_starts -s 7 -c 2 -o gpurun_out/k3_prof_d python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_orfs -s 2 -c 2 -o gpurun_out/orf_prof_d python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
