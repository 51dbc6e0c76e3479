# round 2, GPU call H: parent arrays / dynamic smem / lazy buckets; K1 (4,768) timing; e2e breakdown
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 900 -k "mg_ or config or dropin or fasta or score_all_frames or g3_" --durations=5 ) > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2h_tests.log
for wl in reads400 reads100 contig5m; do
 ( timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2h_$wl.json 2> gpurun_out/r2h_$wl.err; echo "$wl rc=$?"; tail -c 600 gpurun_out/r2h_$wl.err
done
GMG_K1_U=4 GMG_K1_NT=768 timeout 600 python bench.py --workload contig5m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_contig5m_k1_4x768.json 2>/dev/null
GMG_K1_U=4 GMG_K1_NT=512 timeout 600 python bench.py --workload contig5m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_contig5m_k1_4x512.json 2>/dev/null
python - <<'PY'
import json
for f in ('r2h_reads400','r2h_reads100','r2h_contig5m','r2h_contig5m_k1_4x768','r2h_contig5m_k1_4x512'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),'e2e_ms',round(x['e2e'].get('ms_per_step',0),3),'d2h',x['e2e']['d2h_bytes_per_step'],'k',x['roofline'].get('ms_per_step_by_kernel'),'k1ms',x['roofline'].get('kernel_ms'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
python tools/gpu/e2e_breakdown_reads.py 2>&1 | tail -12
