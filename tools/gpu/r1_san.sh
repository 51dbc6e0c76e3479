mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/san.log 2>&1
grep -v "^$" gpurun_out/san.log | head -30 | cut -c1-300
bash tools/gpu/r1_l.sh
