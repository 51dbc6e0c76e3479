# round 2, GPU call S (2 GPUs): NCCL parity test + default bench at N=2 (reads100 with the nested lines) + train500m at N=2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 800 ) > gpurun_out/r2s_tests_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/r2s_tests_multi.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r2s_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2s_bench_n2.json') if l.startswith('{')][-1])
    def show(tag,x):
        print(tag,'n',x['n_gpus'],'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),'parity',x.get('parity_checked'), x['config'].get('sharding'))
    show('reads100',d)
    for k,v in d.get('extra',{}).items(): show(k,v)
except Exception as e: print('no json',e)
PY
