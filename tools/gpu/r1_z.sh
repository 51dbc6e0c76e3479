timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5
python - <<'PY'
import sys, time, subprocess, os
sys.path.insert(0,'tools'); import workloads as W
a, off = W.coding(100100, 333, 7)
W.write_fasta('/tmp/t100m.fa', a, off, prefix='g')
t=time.time(); r=subprocess.run(['glimmer_mg_b200/host/bin/build-icm','-r','/tmp/m_gpu.icm'],stdin=open('/tmp/t100m.fa','rb'),capture_output=True,text=True); print('gpu build-icm 100 Mbp', r.returncode, round(time.time()-t,2),'s', r.stderr[-200:])
a, off = W.coding(2500, 333, 7); W.write_fasta('/tmp/t2m.fa', a, off, prefix='g')
t=time.time(); r=subprocess.run(['oracle/_ref/bin/build-icm','-r','/tmp/m_ref.icm'],stdin=open('/tmp/t2m.fa','rb'),capture_output=True,text=True); print('ref build-icm 2.5 Mbp', r.returncode, round(time.time()-t,2),'s')
t=time.time(); r=subprocess.run(['glimmer_mg_b200/host/bin/build-icm','-r','/tmp/m_gpu2.icm'],stdin=open('/tmp/t2m.fa','rb'),capture_output=True,text=True); print('gpu build-icm 2.5 Mbp', r.returncode, round(time.time()-t,2),'s', open('/tmp/m_gpu2.icm','rb').read()==open('/tmp/m_ref.icm','rb').read())
PY
