# round 2, GPU call O: lanes kernel with per-ORF certificate, ORF finder tweaks: GPU suite, reads100 / contig5m bench, launch list
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=5 ) > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2o_tests.log
( GMG_ORF_TWO_PASS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 900 -k "orfs or config3 or config5" ) > gpurun_out/r2o_tests_twopass.log 2>&1; echo "two-pass rc=$?"; tail -3 gpurun_out/r2o_tests_twopass.log
( GMG_PLAIN_LANES=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 900 -k "plain or config5" ) > gpurun_out/r2o_tests_nolanes.log 2>&1; echo "no-lanes rc=$?"; tail -3 gpurun_out/r2o_tests_nolanes.log
( GMG_MG_FORCE_UNCERT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 900 -k "plain or config5" ) > gpurun_out/r2o_tests_uncert.log 2>&1; echo "force-uncert rc=$?"; tail -3 gpurun_out/r2o_tests_uncert.log
for wl in reads100 contig5m; do
( timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2o_$wl.json 2> gpurun_out/r2o_$wl.err; echo "$wl rc=$?"; tail -c 300 gpurun_out/r2o_$wl.err
done
B="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2o_launches_reads100.csv python bench.py --workload reads100 $B > gpurun_out/r2o_ncu_reads100.log 2>&1
python tools/launch_summary.py gpurun_out/r2o_launches_reads100.csv > gpurun_out/r2o_launch_summary_reads100.txt 2>&1; head -16 gpurun_out/r2o_launch_summary_reads100.txt
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2o_$1 python bench.py --workload $2 $B > gpurun_out/r2o_cap_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2o_$1.ncu-rep --source 14 > gpurun_out/r2o_$1_ncu_full.txt 2>&1; head -52 gpurun_out/r2o_$1_ncu_full.txt
}
cap k_orfs reads100 '^k_orfs$' 3
cap k3_mg_plain_lanes reads100 '^k3_mg_plain_lanes$' 3
rm -f gpurun_out/r2o_*.ncu-rep
python tools/gpu/e2e_breakdown_reads.py 2>&1 | tail -11
python - <<'PY'
import json
for f in ('r2o_reads100','r2o_contig5m'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),x['e2e'].get('ms_per_step'),'k',x['roofline'].get('ms_per_step_by_kernel'), x['roofline'].get('kernel_ms'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
