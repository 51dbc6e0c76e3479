# round 2, GPU call I: thread-per-ORF fused plain kernel, pipelined e2e, (4,768) K1 stderr
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x --timeout 900 -k "mg_ or config5 or config3" --durations=5 ) > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2i_tests.log
for wl in reads100 reads400; do
 ( timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2i_$wl.json 2> gpurun_out/r2i_$wl.err; echo "$wl rc=$?"; tail -c 800 gpurun_out/r2i_$wl.err
done
GMG_PLAIN_WARP=1 timeout 600 python bench.py --workload reads100 --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline > gpurun_out/r2i_reads100_warp.json 2>/dev/null
GMG_K1_U=4 GMG_K1_NT=768 timeout 600 python bench.py --workload contig5m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_k1_4x768.json 2> gpurun_out/r2i_k1_4x768.err; echo "4x768 rc=$?"; tail -c 1500 gpurun_out/r2i_k1_4x768.err
python - <<'PY'
import json
for f in ('r2i_reads100','r2i_reads100_warp','r2i_reads400','r2i_k1_4x768'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',x['e2e'],'k',x['roofline'].get('ms_per_step_by_kernel'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
