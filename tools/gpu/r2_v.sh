# round 2, GPU call V (8 GPUs): reads100 at N=8 with per-rank times (clusters dealt in balanced pairs, adaptive lanes); topology
mkdir -p gpurun_out
nproc > gpurun_out/r2v_topo.txt; nvidia-smi topo -m >> gpurun_out/r2v_topo.txt 2>&1; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2v_topo.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29748 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err; echo "bench n8 rc=$?"; tail -c 300 gpurun_out/r2v_bench_n8.err
( timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra ) > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in (1,8):
  try:
    d=json.loads([l for l in open(f'gpurun_out/r2v_bench_n{n}.json') if l.startswith('{')][-1])
    print('n',n,'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),round(d['e2e']['ms_per_step'],3),'lanes',d['e2e'].get('in_flight'),'one',d['e2e']['one_batch_at_a_time'])
    for r in d['per_rank_ms']['rows']: print('   ',r)
  except Exception as e: print(n,'no json',e)
PY
cat gpurun_out/r2v_topo.txt | head -30
