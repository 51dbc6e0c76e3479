# round 2, GPU call P: K1 second phase (partial windows from the bitmap), extended route tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=5 ) > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2p_tests.log
for wl in reads100 reads400; do
( timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2p_$wl.json 2> gpurun_out/r2p_$wl.err; echo "$wl rc=$?"; tail -c 300 gpurun_out/r2p_$wl.err
done
( GMG_K1_FIX=3 timeout 600 python bench.py --workload reads100 --steps 20 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2p_reads100_fix3.json 2> gpurun_out/r2p_reads100_fix3.err; echo "fix3 rc=$?"
B="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2p_launches_reads100.csv python bench.py --workload reads100 $B > gpurun_out/r2p_ncu_reads100.log 2>&1
python tools/launch_summary.py gpurun_out/r2p_launches_reads100.csv > gpurun_out/r2p_launch_summary_reads100.txt 2>&1; head -12 gpurun_out/r2p_launch_summary_reads100.txt
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2p_$1 python bench.py --workload $2 $B > gpurun_out/r2p_cap_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2p_$1.ncu-rep --source 45 > gpurun_out/r2p_$1_ncu_full.txt 2>&1; head -24 gpurun_out/r2p_$1_ncu_full.txt | cut -c1-150
}
cap k1_reads reads100 '^k1_planes_bucketed$' 3
cap k3_mg_plain_lanes reads100 '^k3_mg_plain_lanes$' 3
rm -f gpurun_out/r2p_*.ncu-rep
python - <<'PY'
import json
for f in ('r2p_reads100','r2p_reads100_fix3','r2p_reads400'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),x['e2e'].get('ms_per_step'),'k',x['roofline'].get('ms_per_step_by_kernel'), x['roofline'].get('kernel_ms'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
