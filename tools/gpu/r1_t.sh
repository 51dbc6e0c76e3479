timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "single_pass" --tb=short 2>&1 | grep -v "^$" | tail -25 | cut -c1-300
