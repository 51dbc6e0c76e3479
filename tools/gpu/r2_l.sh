# round 2, GPU call L: fresh launch list of reads100 (device-resident + e2e), full captures of the e2e-side kernels,
# GPU core dump of the (4,768) K1 fault, contig5m with the small-batch partial fix
mkdir -p gpurun_out
B="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2l_launches_reads100.csv python bench.py --workload reads100 $B > gpurun_out/r2l_ncu_reads100.log 2>&1
python tools/launch_summary.py gpurun_out/r2l_launches_reads100.csv > gpurun_out/r2l_launch_summary_reads100.txt 2>&1; head -30 gpurun_out/r2l_launch_summary_reads100.txt
cap() { # name workload kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -f -o gpurun_out/r2l_$1 python bench.py --workload $2 $B > gpurun_out/r2l_cap_$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2l_$1.ncu-rep --source 14 > gpurun_out/r2l_$1_ncu_full.txt 2>&1; head -50 gpurun_out/r2l_$1_ncu_full.txt
}
cap k_orfs reads100 'k_orfs<' 3
cap k1_partial_fix reads100 k1_partial_fix 3
cap k3_mg_reduce reads100 k3_mg_reduce 2
cap k_bucket_finish reads100 k_bucket_finish 3
cap k3_mg_plain reads100 'k3_mg_plain\(' 3
rm -f gpurun_out/r2l_*.ncu-rep
( timeout 600 python bench.py --workload contig5m --steps 20 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2l_contig5m.json 2> gpurun_out/r2l_contig5m.err; echo "contig5m rc=$?"
# (4,768): GPU core dump at the exception, read back with cuda-gdb on the box
export GMG_K1_U=4 GMG_K1_NT=768 GMG_G3_NO_SIDE=1
export CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_ENABLE_LIGHTWEIGHT_COREDUMP=1 CUDA_COREDUMP_FILE=/tmp/gmgcore CUDA_COREDUMP_SHOW_PROGRESS=1
timeout 300 python bench.py --workload contig5m --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2l_4x768_core.json 2> gpurun_out/r2l_4x768_core.err; echo "4x768 rc=$?"; tail -c 600 gpurun_out/r2l_4x768_core.err
unset CUDA_ENABLE_COREDUMP_ON_EXCEPTION CUDA_ENABLE_LIGHTWEIGHT_COREDUMP GMG_K1_U GMG_K1_NT GMG_G3_NO_SIDE
ls -la /tmp/gmgcore* 2>&1 | head
for f in /tmp/gmgcore*; do
  timeout 300 cuda-gdb -batch -ex "target cudacore $f" -ex "info cuda kernels" -ex "info cuda lanes" -ex "bt" -ex "info registers pc" -ex "x/8i \$pc-32" -ex "info cuda exception" > gpurun_out/r2l_4x768_cudagdb.txt 2>&1
  break
done
tail -60 gpurun_out/r2l_4x768_cudagdb.txt | cut -c1-250
python - <<'PY'
import json
for f in ('r2l_contig5m',):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',x['e2e']['value'],x['e2e'].get('ms_per_step'), x['roofline'].get('kernel_ms'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
