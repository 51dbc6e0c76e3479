mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -q --timeout 600 2>&1 | tail -15
