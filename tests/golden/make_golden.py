#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the build container).

Two kinds of fixtures:
 * verbatim DATA files from /root/reference/sample-run (inputs, models and the authors'
   golden outputs; no source code), gzip'd where text;
 * outputs of the UNMODIFIED reference compiled by oracle/Makefile into oracle/_ref
   (raw ORF/start-list dumps from the instrumented drivers, `.predict` files), so the
   GPU-side tests can compare against the real implementation without /root/reference.

usage: python tests/golden/make_golden.py
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
S = "/root/reference/sample-run"
REF = os.path.join(ROOT, "oracle", "_ref", "bin")


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def run(cmd, **env):
    subprocess.run(cmd, check=True, env=dict(os.environ, **env), stderr=subprocess.DEVNULL)


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    # --- verbatim data ---
    gz_copy(f"{S}/glimmer-mg/seqs.fa", f"{HERE}/seqs.fa.gz")
    for k in (4, 5):
        shutil.copy(f"{S}/glimmer-mg/results/cluster-{k}.icm", f"{HERE}/cluster-{k}.icm")
        shutil.copy(f"{S}/glimmer-mg/results/icm-{k}.scores.tmp", f"{HERE}/icm-{k}.scores.tmp")
    shutil.copy(f"{S}/glimmer-mg/results/seqs.cluster-4.run1.filt.gicm", HERE)
    gz_copy(f"{S}/glimmer-mg/results/seqs.cluster-4.run1.filt.gene.fasta", f"{HERE}/seqs.cluster-4.run1.filt.gene.fasta.gz")
    gz_copy(f"{S}/glimmer-mg/results/seqs.cluster-5.run1.filt.gene.fasta", f"{HERE}/seqs.cluster-5.run1.filt.gene.fasta.gz")
    shutil.copy(f"{S}/glimmer-mg/results/seqs.cluster-5.run1.filt.gicm", HERE)
    gz_copy(f"{S}/glimmer3/NC_000915.fna", f"{HERE}/NC_000915.fna.gz")
    gz_copy(f"{S}/glimmer3/results/NC_000915.train", f"{HERE}/NC_000915.train.gz")
    shutil.copy(f"{S}/glimmer3/results/NC_000915.icm", HERE)
    gz_copy(f"{S}/glimmer3/results/NC_000915.run1.predict", f"{HERE}/NC_000915.run1.predict.gz")
    # --- reference-generated vectors ---
    tmp = "/tmp/make_golden"
    os.makedirs(tmp, exist_ok=True)
    icm = f"{S}/glimmer3/results/NC_000915.icm"
    recs = []
    with open(f"{S}/glimmer-mg/seqs.fa") as f:
        for line in f:
            if line.startswith(">"):
                recs.append([line, ""])
            else:
                recs[-1][1] += line
    for tag, n, flags in (("plain", 120, []), ("indel", 40, ["-i"]), ("sub", 80, ["-s"])):
        fa = f"{tmp}/reads_{tag}.fa"
        with open(fa, "w") as f:
            for h, s in recs[:n]:
                f.write(h + s)
        dump = f"{tmp}/mg_{tag}.dump"
        run([f"{REF}/glimmer-mg-dump", "-u", "1.0", "-m", icm] + flags + [fa, f"{tmp}/mg_{tag}"],
            GMG_DUMP=dump, GMG_DUMP_FS="25" if tag == "plain" else "0")
        gz_copy(dump, f"{HERE}/mg_{tag}_{n}.dump.gz")
        gz_copy(f"{tmp}/mg_{tag}.predict", f"{HERE}/mg_{tag}_{n}.predict.gz")
    # glimmer3 raw start lists for the first 300 kbp of NC_000915
    fna = f"{tmp}/nc300k.fna"
    with open(f"{S}/glimmer3/NC_000915.fna") as f, open(fna, "w") as g:
        lines = f.readlines()
        g.write(lines[0])
        g.writelines(lines[1:1 + 300000 // 70])
    run([f"{REF}/glimmer3-dump", "-u", "-12", "-m", icm, fna, f"{tmp}/g3_300k"], GMG_DUMP=f"{tmp}/g3_300k.dump")
    gz_copy(f"{tmp}/g3_300k.dump", f"{HERE}/g3_300k.dump.gz")
    gz_copy(f"{tmp}/g3_300k.predict", f"{HERE}/g3_300k.predict.gz")
    print("golden fixtures written to", HERE)
    for fn in sorted(os.listdir(HERE)):
        print("%10d  %s" % (os.path.getsize(os.path.join(HERE, fn)), fn))


if __name__ == "__main__":
    sys.exit(main())
