timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "score_all_frames or many_contigs or g3_full_genome" 2>&1 | grep "=========\|passed\|failed" | head -6
