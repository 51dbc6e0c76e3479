mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "mg_start_lists or fasta or histogram or many_models or two_pass or find_orfs_reads" 2>&1 | grep -v "^$" | tail -25 | cut -c1-300
