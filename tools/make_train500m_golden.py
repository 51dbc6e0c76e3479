"""One-off generator of tests/golden/train500m.sha256: the model file the UNMODIFIED reference build-icm (oracle/_ref,
built from /root/reference by oracle/Makefile) writes for BASELINE.json configs[3] at full size (500 500 x 999 bp
stop-free coding strings, seed 7, `build-icm -r`).  About 25 minutes on one core; run in the build container.
bench.py --workload train500m compares the device-trained model file against this digest (parity_checked)."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads as W  # noqa: E402

n_seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_500
tmp = sys.argv[2] if len(sys.argv) > 2 else "/tmp/t500"
os.makedirs(tmp, exist_ok=True)
a, off = W.coding(n_seqs, 333, W.TRAIN_SEED)
fa = os.path.join(tmp, "train.fa")
W.write_fasta(fa, a, off, prefix="g")
out = os.path.join(tmp, "ref.icm")
t0 = time.time()
with open(fa, "rb") as fin:
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "bin", "build-icm"), "-r", out], stdin=fin, check=True)
sec = time.time() - t0
with open(out, "rb") as fp:
    digest = hashlib.sha256(fp.read()).hexdigest()
rec = {"n_seqs": n_seqs, "codons": 333, "seed": W.TRAIN_SEED, "bases": int(off[-1]), "command": "build-icm -r",
       "model_file_sha256": digest, "model_file_bytes": os.path.getsize(out), "reference_seconds_one_core": round(sec, 1)}
name = "train500m.sha256.json" if n_seqs == 500_500 else f"train_{n_seqs}.sha256.json"
with open(os.path.join(ROOT, "tests", "golden", name), "w") as fp:
    json.dump(rec, fp, indent=1)
print(rec)
