"""Parity checks on the synthetic BASELINE.json configs (tools/workloads.py): the checker half.

TEST INFRASTRUCTURE: compares results of the CUDA path (obtained by the caller through the C-ABI) with the
oracle port (oracle/icm_oracle.c via tests/oracle_lib.py) and, where oracle/_ref was built, with the
unmodified reference binaries.  Used by tests/test_gpu_configs.py and by bench.py's `parity_checked` leg
(outside the timed region).  Nothing in glimmer_mg_b200/ imports this module.

BASELINE.md section 3: every timing on a synthetic config is accompanied by a check of its output against
the oracle; the bar here is 100 % identical (ORF tables, start lists incl. FP64 score bits, model files).
"""
import hashlib
import os
import subprocess
import tempfile

import numpy as np

import oracle_lib as O

ROOT = O.ROOT
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")


class ParityError(AssertionError):
    pass


def _require(cond, msg):
    if not cond:
        raise ParityError(msg)


def oracle_model(gpu_model=None, path=None):
    """orc_icm* of a model file, or of a device model (written with gmg_icm_write first)."""
    if path is None:
        fd, path = tempfile.mkstemp(suffix=".icm", prefix="gmg_parity_")
        os.close(fd)
        try:
            gpu_model.Output(path)
            return O.lib().orc_icm_read(path.encode())
        finally:
            os.unlink(path)
    return O.lib().orc_icm_read(os.fsencode(path))


def check_scoring(kind, ascii_arr, off, seq_ids, orfs, ooff, starts, soff, ogene, gc, stops, **pkw):
    """Sequences `seq_ids` of a batch: GPU ORF table + start lists against the oracle port.

    kind   'g3' (glimmer3 Score_Orfs) or 'mg' (glimmer-mg Score_Orfs_Errors)
    ascii_arr, off   the batch as given to gmg_seqset_create
    orfs, ooff, starts, soff   what gmg_get_orfs / gmg_get_starts returned for the whole batch
    ogene  orc_icm* of the gene model; the independent model is rebuilt from (gc, stops)
    pkw    oracle parameter overrides (allow_indels=1, ignore_score_len=...)
    -> dict(seqs, orfs, starts) counted over the checked sequences; raises ParityError on any difference."""
    oi = O.build_indep(gc, stops)
    op = O.params(kind == "mg", **pkw)
    raw = ascii_arr.tobytes() if hasattr(ascii_arr, "tobytes") else bytes(ascii_arr)
    n_orfs = n_starts = 0
    for i in seq_ids:
        s = O.filter_lower(raw[int(off[i]):int(off[i + 1])])
        worfs = O.find_orfs(s, op)
        got_orfs = orfs[int(ooff[i]):int(ooff[i + 1])]
        _require(got_orfs.tobytes() == worfs.tobytes(),
                 f"{kind}: ORF table of sequence {i} differs from the oracle ({len(got_orfs)} vs {len(worfs)} ORFs)")
        woff, wst = (O.mg_score_orfs if kind == "mg" else O.g3_score_orfs)(ogene, oi, s, op, worfs)
        a, b = int(soff[int(ooff[i])]), int(soff[int(ooff[i + 1])])
        got = starts[a:b]
        _require(len(got) == len(wst), f"{kind}: sequence {i}: {len(got)} starts, oracle {len(wst)}")
        _require(got.tobytes() == wst.tobytes(), f"{kind}: start lists of sequence {i} differ from the oracle "
                 f"(first difference at record {_first_diff(got, wst)})")
        rel = np.asarray(soff[int(ooff[i]):int(ooff[i + 1]) + 1], np.int64) - a
        _require((rel == np.asarray(woff, np.int64)).all(), f"{kind}: per-ORF start offsets of sequence {i} differ")
        n_orfs += len(worfs)
        n_starts += len(wst)
    return {"seqs": len(seq_ids), "orfs": n_orfs, "starts": n_starts}


def _first_diff(a, b):
    for k in range(min(len(a), len(b))):
        if a[k].tobytes() != b[k].tobytes():
            return f"{k}: got {a[k]} want {b[k]}"
    return "length"


def sample_ids(n, want, seed=1):
    """`want` sequence indices of a batch of n: the first and last few plus a seeded random sample."""
    if n <= want:
        return list(range(n))
    edge = min(8, want // 4)
    rng = np.random.default_rng(seed)
    mid = rng.choice(np.arange(edge, n - edge), size=want - 2 * edge, replace=False)
    return sorted(set(range(edge)) | set(range(n - edge, n)) | set(int(x) for x in mid))


def file_sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as fp:
        for blk in iter(lambda: fp.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def have_ref_bin(name):
    return os.path.exists(os.path.join(REFBIN, name))


def run(cmd, stdin_path=None, timeout=3600):
    fin = open(stdin_path, "rb") if stdin_path else None
    try:
        r = subprocess.run(cmd, stdin=fin, capture_output=True, text=True, timeout=timeout)
    finally:
        if fin:
            fin.close()
    _require(r.returncode == 0, f"{' '.join(cmd)} failed:\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}")
    return r


def predict_pair(driver, args, fasta, workdir, tag="p"):
    """Run oracle/_ref/bin/<driver> and <driver>-gmg (the reference driver compiled against the C-ABI) with the
    same arguments -> (reference .predict bytes, GPU-path .predict bytes)."""
    out = []
    for exe, name in ((driver, tag + "_ref"), (driver + "-gmg", tag + "_gmg")):
        run([os.path.join(REFBIN, exe), *args, fasta, os.path.join(workdir, name)])
        with open(os.path.join(workdir, name + ".predict"), "rb") as fp:
            out.append(fp.read())
    return out[0], out[1]


def predict_identity(ref_bytes, got_bytes):
    """Fraction of the reference's .predict lines (headers included) found, in order, in the GPU path's."""
    a, b = ref_bytes.split(b"\n"), got_bytes.split(b"\n")
    if a == b:
        return 1.0
    import difflib
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    same = sum(m.size for m in sm.get_matching_blocks())
    return same / max(1, len(a))
