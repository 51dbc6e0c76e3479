timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "score_string" 2>&1 | tail -12
python - <<'PY'
import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import numpy as np, glimmer_mg_b200 as g, workloads as W, oracle_lib as O
ctx = g.Context(0)
names = ["cluster-4.icm", "cluster-5.icm", "NC_000915.icm"]
models = [g.ICM.Read(ctx, os.path.join('tests/golden', nm)) for nm in names] * 6
c = W.contig(W.CONTIG_SEED, 5_000_000)
a, off = W.reads(c, 250000, 400, 11, indel=False)
ss = g.SeqSet(ctx, ascii=a, offsets=off)
for rep in range(3):
    t=time.perf_counter(); out = g.score_strings_many(ctx, models, ss, 0); dt=time.perf_counter()-t
print(f"{len(models)} models x {ss.n} reads x 400 bp: {dt*1e3:.1f} ms -> {len(models)*ss.total/dt/1e9:.1f} G model-bases/s")
t=time.perf_counter()
for m in models: m.score_strings(ss, 0)
print('per-model ordered calls', round((time.perf_counter()-t)*1e3,1), 'ms')
om = O.lib().orc_icm_read(b'tests/golden/cluster-4.icm'); s_all=a.tobytes()
t=time.perf_counter()
for i in range(5000): O.lib().orc_score_string(om, s_all[off[i]:off[i+1]], 400, 0)
dt=time.perf_counter()-t; print('oracle port 1 core', 5000*400/dt/1e9, 'G model-bases/s')
PY
