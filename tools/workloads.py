"""Synthetic workloads of BASELINE.json's configs, shared by bench.py and the
full-size property tests.  Thin ctypes wrapper over tools/libgmgsynth.so (tools/synth.c, xoshiro256**).

Nothing here touches /root/reference: the codon table comes from the committed fixture
tests/golden/NC_000915.train.gz and the gene model from tests/golden/NC_000915.icm.
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_SO = os.path.join(ROOT, "tools", "libgmgsynth.so")
_lib = None

CONTIG_SEED = 20261017   # config 2
READS400_SEED = 42       # config 3
TRAIN_SEED = 7           # config 4
READS100_SEED = 5        # config 5


def build():
    src = os.path.join(ROOT, "tools", "synth.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", _SO, src, "-lm"], check=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.synth_contig.restype = C.c_int64
        L.synth_contig.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_double, C.c_void_p]
        L.synth_reads.restype = C.c_int64
        L.synth_reads.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.synth_coding.restype = C.c_int64
        L.synth_coding.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def gene_model_path():
    return os.path.join(GOLDEN, "NC_000915.icm")


_codon_freq = None


def codon_freq():
    """Codon usage of the sample genome's training genes (64 doubles, index 16*b0 + 4*b1 + b2, acgt = 0..3)."""
    global _codon_freq
    if _codon_freq is None:
        code = np.full(256, 255, np.uint8)
        for i, ch in enumerate(b"acgt"):
            code[ch] = i
            code[ch - 32] = i
        freq = np.ones(64, np.float64)  # +1 smoothing
        seq = []
        with gzip.open(os.path.join(GOLDEN, "NC_000915.train.gz"), "rb") as fp:
            for line in fp:
                if line.startswith(b">"):
                    if seq:
                        freq += _count(code, b"".join(seq))
                    seq = []
                else:
                    seq.append(line.strip())
        if seq:
            freq += _count(code, b"".join(seq))
        _codon_freq = freq / freq.sum()
    return _codon_freq


def _count(code, s):
    v = code[np.frombuffer(s, np.uint8)]
    n = len(v) // 3 * 3
    v = v[:n].reshape(-1, 3).astype(np.int64)
    ok = (v < 4).all(axis=1)
    idx = (v[:, 0] * 16 + v[:, 1] * 4 + v[:, 2])[ok]
    return np.bincount(idx, minlength=64).astype(np.float64)


def reweight_gc(freq, gc):
    """Codon table re-weighted towards a target GC fraction (config 5's per-genome tables)."""
    w = np.zeros(64)
    for c in range(64):
        ngc = sum(1 for b in ((c >> 4) & 3, (c >> 2) & 3, c & 3) if b in (1, 2))
        w[c] = freq[c] * (gc ** ngc) * ((1.0 - gc) ** (3 - ngc))
    return w / w.sum()


def contig(seed=CONTIG_SEED, length=5_000_000, freq=None, gc=0.39, out=None):
    """Config 2: one synthetic bacterial contig -> uint8 array of lower-case acgt."""
    freq = np.ascontiguousarray(codon_freq() if freq is None else freq, np.float64)
    buf = np.empty(length + 64, np.uint8) if out is None else out
    n = lib().synth_contig(seed, length, freq.ctypes.data, gc, buf.ctypes.data)
    return buf[:n]


def reads(contig_arr, n_reads, read_len, seed, indel=False):
    """Configs 3/5: reads drawn from a contig -> (uint8 ascii, int64 offsets)."""
    contig_arr = np.ascontiguousarray(contig_arr, np.uint8)
    out = np.empty(n_reads * (read_len + read_len // 3 + 2), np.uint8)
    off = np.empty(n_reads + 1, np.int64)
    n = lib().synth_reads(seed, contig_arr.ctypes.data, len(contig_arr), n_reads, read_len, 1 if indel else 0,
                          out.ctypes.data, off.ctypes.data)
    return out[:n], off


def coding(n_seqs, codons=333, seed=TRAIN_SEED, freq=None):
    """Config 4: stop-free coding sequences -> (uint8 ascii, int64 offsets)."""
    freq = np.ascontiguousarray(codon_freq() if freq is None else freq, np.float64)
    out = np.empty(n_seqs * 3 * codons, np.uint8)
    n = lib().synth_coding(seed, n_seqs, codons, freq.ctypes.data, out.ctypes.data)
    off = np.arange(n_seqs + 1, dtype=np.int64) * (3 * codons)
    return out[:n], off


def write_fasta(path, ascii_arr, off=None, prefix="seq", width=60):
    """Lower-case multi-FASTA, 60 columns (what the reference binaries read)."""
    if off is None:
        off = np.array([0, len(ascii_arr)], np.int64)
    with open(path, "wb") as fp:
        for i in range(len(off) - 1):
            s = ascii_arr[off[i]:off[i + 1]]
            fp.write(b">%s%d\n" % (prefix.encode(), i))
            n = len(s)
            full = n // width * width
            if full:
                body = np.empty((full // width, width + 1), np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                fp.write(body.tobytes())
            if n > full:
                fp.write(s[full:].tobytes() + b"\n")
