# final evidence of a round: launch lists of the four workloads, ncu --set full captures of the hot kernels,
# compute-sanitizer memcheck / racecheck over their tests (outputs under gpurun_out/, copy what is to be judged to profiles/)
mkdir -p gpurun_out
B="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
for wl in reads100 reads400 contig5m train500m; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_final_launches_$wl.csv python bench.py --workload $wl $B > gpurun_out/r2_final_ncu_$wl.log 2>&1
  python tools/launch_summary.py gpurun_out/r2_final_launches_$wl.csv > gpurun_out/r2_final_launch_summary_$wl.txt 2>&1; head -14 gpurun_out/r2_final_launch_summary_$wl.txt | cut -c1-130
done
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2_final_$1 python bench.py --workload $2 $B > gpurun_out/r2_final_cap_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_final_$1.ncu-rep --source 16 > gpurun_out/r2_final_$1_ncu_full.txt 2>&1; head -8 gpurun_out/r2_final_$1_ncu_full.txt | cut -c1-150
}
cap k1_planes_bucketed_reads100 reads100 '^k1_planes_bucketed$' 3
cap k1_partial_fix reads100 '^k1_partial_fix$' 3
cap k3_mg_plain_lanes reads100 '^k3_mg_plain_lanes$' 3
cap k_orfs reads100 '^k_orfs$' 3
cap k_bucket_finish reads100 '^k_bucket_finish$' 3
cap k3_mg_reduce_small reads100 '^k3_mg_reduce_small$' 2
cap k1_planes_bucketed_contig5m contig5m '^k1_planes_bucketed$' 3
rm -f gpurun_out/r2_final_*.ncu-rep
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1100 -k "find_orfs or plain_fused or reduction or quality_ingest or score_all_frames_bit_exact" > gpurun_out/r2_final_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_final_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 1100 -k "find_orfs_reads or plain_fused or quality_ingest" > gpurun_out/r2_final_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_final_sanitizer_racecheck.log
