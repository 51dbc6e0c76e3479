timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "mg_start" --tb=short 2>&1 | grep -v "^$" | head -60 | cut -c1-600
