timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2.out 2> gpurun_out/n2.err
echo "stdout lines: $(wc -l < gpurun_out/n2.out)"; cut -c1-200 gpurun_out/n2.out; grep -c "NCCL version" gpurun_out/n2.err
timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -1
