# round 2, GPU call U (8 GPUs): the default bench at N=8 under torchrun (what the driver's scaling run does), then N=4
mkdir -p gpurun_out
for n in 8 4; do
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2u_bench_n$n.json 2> gpurun_out/r2u_bench_n$n.err; echo "bench n$n rc=$?"; tail -c 400 gpurun_out/r2u_bench_n$n.err
done
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in (1,4,8):
  try:
    d=json.loads([l for l in open(f'gpurun_out/r2u_bench_n{n}.json') if l.startswith('{')][-1])
    def show(tag,x):
        print(tag,'n',x['n_gpus'],'value',round(x['value'],2),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value'],2),round(x['e2e'].get('ms_per_step',0),3),'parity',x.get('parity_checked'),(x['roofline'].get('ms_per_step_by_kernel') or x['roofline'].get('kernel_ms')))
    show('reads100',d)
    for k,v in d.get('extra',{}).items(): show(k,v)
  except Exception as e: print(n,'no json',e)
PY
