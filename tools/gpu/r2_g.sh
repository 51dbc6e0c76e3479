# round 2, GPU call G: launch lists + ncu --set full captures of the current kernels, sanitizer logs
mkdir -p gpurun_out
B="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
for wl in reads100 reads400 contig5m train500m; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workload $wl $B > gpurun_out/r2g_ncu_$wl.log 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_$wl.csv > gpurun_out/r2_launch_summary_$wl.txt 2>&1; head -24 gpurun_out/r2_launch_summary_$wl.txt
done
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2_$1 python bench.py --workload $2 $B > gpurun_out/r2g_cap_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_$1.ncu-rep --source 12 > gpurun_out/r2_$1_ncu_full.txt 2>&1; head -40 gpurun_out/r2_$1_ncu_full.txt
}
cap k1_planes_bucketed contig5m k1_planes_bucketed 3
cap k2_g3_codon_cum contig5m k2_g3_codon_cum 3
cap k2_prefix_lanes reads400 k2_prefix_lanes 3
cap k_mgf_c reads400 k_mgf_c 3
cap k_mgf_w2 reads400 k_mgf_w2 3
cap k_mgf_b reads400 k_mgf_b 3
cap k3_mg_reduce reads400 k3_mg_reduce 3
cap k3_mg_plain reads100 'k3_mg_plain\(' 3
cap k4_hist_level train500m k4_hist_level 9
cap k4_hist_build train500m k4_hist_build 1
# sanitizers: memcheck over the new kernels' tests, racecheck over K4 (shared-memory / global atomics)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1400 -k "mg_flat or reduction or plain_fused or all_frame or level_finish" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1400 -k "training_histogram or count_level" > gpurun_out/r2_sanitizer_racecheck_k4.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_racecheck_k4.log
# the (4, 768) K1 shape that round 1 reported as faulting
GMG_K1_U=4 GMG_K1_NT=768 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python bench.py --workload contig5m --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_k1_4x768_memcheck.log 2>&1; echo "k1 4x768 rc=$?"; tail -12 gpurun_out/r2_k1_4x768_memcheck.log
ls -la gpurun_out | grep r2_ | head -50
