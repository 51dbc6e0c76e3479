# round 2, GPU call J: staged ORF finder; (4,768) K1 fault hunt
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 900 -k "orfs or mg_plain or config or dropin or g3_" --durations=5 ) > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2j_tests.log
( timeout 600 python bench.py --workload reads100 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2j_reads100.json 2> gpurun_out/r2j_reads100.err; echo "reads100 rc=$?"; tail -c 500 gpurun_out/r2j_reads100.err
python tools/gpu/e2e_breakdown_reads.py 2>&1 | tail -11
export GMG_K1_U=4 GMG_K1_NT=768
CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --workload contig5m --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2j_4x768_blocking.json 2> gpurun_out/r2j_4x768_blocking.err; echo "blocking rc=$?"; tail -c 700 gpurun_out/r2j_4x768_blocking.err
GMG_G3_NO_SIDE=1 timeout 300 python bench.py --workload contig5m --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2j_4x768_noside.json 2> gpurun_out/r2j_4x768_noside.err; echo "noside rc=$?"; tail -c 700 gpurun_out/r2j_4x768_noside.err
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python bench.py --workload contig5m --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-parity > gpurun_out/r2j_4x768_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -v "^{" gpurun_out/r2j_4x768_racecheck.log | tail -15 | cut -c1-220
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 python bench.py --workload contig5m --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-parity > gpurun_out/r2j_4x768_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -v "^{" gpurun_out/r2j_4x768_initcheck.log | tail -15 | cut -c1-220
unset GMG_K1_U GMG_K1_NT
python - <<'PY'
import json
for f in ('r2j_reads100','r2j_4x768_blocking','r2j_4x768_noside'):
    try:
        x=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',x['e2e']['value'],x['e2e'].get('ms_per_step'),'k',x['roofline'].get('ms_per_step_by_kernel'), x['roofline'].get('kernel_ms'),'parity',x.get('parity_checked'))
    except Exception as e: print(f,'no json',e)
PY
