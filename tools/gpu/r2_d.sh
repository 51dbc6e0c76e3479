# round 2, GPU call D (2 GPUs): training with device-side level finish + cell-sharded histogram; NCCL parity; bench N=1/N=2
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 900 -k "training or train or count_level or multi or build_icm or config4 or config5" --durations=6 ) > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2d_tests.log
( timeout 600 python bench.py --workload train500m --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2d_train_n1.json 2> gpurun_out/r2d_train_n1.err; echo "bench n1 rc=$?"; tail -c 800 gpurun_out/r2d_train_n1.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train500m --steps 10 --warmup 3 ) > gpurun_out/r2d_train_n2.json 2> gpurun_out/r2d_train_n2.err; echo "bench n2 rc=$?"; tail -c 800 gpurun_out/r2d_train_n2.err
python - <<'PY'
import json
for f in ('r2d_train_n1','r2d_train_n2'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f,'value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],2),'k4 share',round(d['roofline']['kernel_share_of_step'],3),'flagged',d['config'].get('host_recomputed_nodes'),'parity',d['parity'])
    except Exception as e: print(f,'no json',e)
PY
