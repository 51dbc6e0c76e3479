"""The synthetic workload generator (tools/synth.c) is deterministic and produces what SURVEY.md 8(d) specifies."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import workloads as W


def test_contig_is_deterministic_and_lowercase_acgt():
    a = W.contig(length=200_000)
    b = W.contig(length=200_000)
    c = W.contig(seed=W.CONTIG_SEED + 1, length=200_000)
    assert len(a) == 200_000 and (a == b).all() and not (a == c).all()
    assert set(np.unique(a).tolist()) <= set(b"acgt")
    gc = np.isin(a, list(b"cg")).mean()
    assert 0.3 < gc < 0.5


def test_reads_lengths_and_indels():
    c = W.contig(length=100_000)
    r, off = W.reads(c, 2000, 100, W.READS100_SEED, indel=False)
    assert (np.diff(off) == 100).all() and len(r) == off[-1]
    r2, off2 = W.reads(c, 2000, 400, W.READS400_SEED, indel=True)
    d = np.diff(off2)
    assert d.min() >= 380 and d.max() <= 420 and (d != 400).any()
    # error-free forward-strand reads are substrings of the contig
    s = c.tobytes()
    hits = sum(1 for i in range(50) if r[off[i]:off[i + 1]].tobytes() in s)
    assert 10 <= hits <= 45  # about half are reverse-complemented


def test_coding_sequences_are_stop_free():
    s, off = W.coding(300, 333)
    assert (np.diff(off) == 999).all()
    cod = s.reshape(-1, 3)
    stops = {b"taa", b"tag", b"tga"}
    assert not any(bytes(x) in stops for x in cod[:20000])


def test_fasta_writer_roundtrip(tmp_path):
    c = W.contig(length=1234)
    p = str(tmp_path / "x.fa")
    W.write_fasta(p, c, prefix="contig")
    lines = open(p, "rb").read().split(b"\n")
    assert lines[0] == b">contig0" and all(len(l) == 60 for l in lines[1:-2])
    assert b"".join(lines[1:]) == c.tobytes()
