for r in 4 8 16 0; do
GMG_K1_R=$r timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "R=$r rc=$? illegal=$(grep -c illegal gpurun_out/bench_v.err) $(python -c "import json;d=json.load(open('gpurun_out/bench_v.json'));print(round(d['roofline']['kernel_ms']*1e3,1), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))" 2>/dev/null)"
GMG_K1_R=$r timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "score_all_frames or g3_start_lists or many_contigs" 2>&1 | tail -1
done
