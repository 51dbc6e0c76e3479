mkdir -p gpurun_out
for w in contig5m reads100 train500m; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_n8.json 2> gpurun_out/bench_${w}_n8.err
tail -2 gpurun_out/bench_${w}_n8.err | cut -c1-300; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${w}_n8.json')); print('$w', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('fail', e)
PY
done
