# scaling of the default workload (reads100) at N = 8, 4, 2 and of training at N = 8, final code
mkdir -p gpurun_out
for n in 8 4 2; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2975$n bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2z_reads100_n$n.json 2> gpurun_out/r2z_reads100_n$n.err; echo "reads100 n$n rc=$?"; tail -c 200 gpurun_out/r2z_reads100_n$n.err
done
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29759 bench.py --gpus 8 --workload train500m --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2z_train500m_n8.json 2> gpurun_out/r2z_train500m_n8.err; echo "train n8 rc=$?"
( timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2z_reads100_n1.json 2> gpurun_out/r2z_reads100_n1.err; echo "reads100 n1 rc=$?"
python - <<'PY'
import json
for f in ('reads100_n1','reads100_n2','reads100_n4','reads100_n8','train500m_n8'):
  try:
    d=json.loads([l for l in open(f'gpurun_out/r2z_{f}.json') if l.startswith('{')][-1])
    print(f,'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),round(d['e2e']['ms_per_step'],3),'lanes',d['e2e'].get('in_flight'),'one',d['e2e'].get('one_batch_at_a_time'), d.get('parity_checked'))
    for r in d.get('per_rank_ms',{}).get('rows',[]): print('   ',r)
  except Exception as e: print(f,'no json',e)
PY
