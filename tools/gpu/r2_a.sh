# round 2, GPU call A: full GPU test suite (incl. the new config parity gates), L2 peak microbenchmark, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=15 ) > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2a_tests.log
tools/gpu/ubench_l2 > gpurun_out/r2a_l2.json 2> gpurun_out/r2a_l2.err; cat gpurun_out/r2a_l2.json
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.err; head -c 6000 gpurun_out/r2a_bench.json
