timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
python tools/gpu/e2e_breakdown_reads.py reads100 2>&1 | grep "find_orfs\|seqset_create\|score"
python tools/gpu/e2e_breakdown.py 2>&1 | grep "find_orfs\|seqset_create\|score"
