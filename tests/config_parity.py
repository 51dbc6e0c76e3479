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
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")       # the unmodified reference (test infrastructure)
HOSTBIN = os.path.join(ROOT, "glimmer_mg_b200", "host", "bin")  # the product's C++ hosts over the C-ABI


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
    return os.path.exists(os.path.join(HOSTBIN if name.endswith("-gmg") else REFBIN, name))


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
    """Run oracle/_ref/bin/<driver> and glimmer_mg_b200/host/bin/<driver>-gmg (the reference driver compiled against
    the C-ABI) with the same arguments -> (reference .predict bytes, GPU-path .predict bytes)."""
    out = []
    for exe, name in ((os.path.join(REFBIN, driver), tag + "_ref"), (os.path.join(HOSTBIN, driver + "-gmg"), tag + "_gmg")):
        run([exe, *args, fasta, os.path.join(workdir, name)])
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


def reduce_reference(raw, frame, stop_position, seq_len, min_gene_len, model, seq_cls=0):
    """What the reference keeps of one ORF's raw start_list (structured array in generation order): restatement of
    Score_Orfs_Errors' filter (glimmer-mg.cc:1656-1684) and of the per-position arg-max of Add_Events_Fwd / _Rev
    (glimmer_base.cc:65-128, 175-235) WITHOUT an RBS model, event scores added in the reference's order.
    -> ("drop", None) | ("keep", {pos: record index}) | ("undecided", None) where the outcome hangs on std::sort's
    order among equal positions or on an exact score tie (the device must hand such ORFs back with status 2)."""
    n = len(raw)
    if n == 0:
        return "drop", None
    pos = raw["pos"]
    ext = pos.min() if frame > 0 else pos.max()
    js = raw["j"][pos == ext]
    ok = js + 1 >= min_gene_len
    if ok.any() and not ok.all():
        return "undecided", None
    if not ok.all():
        return "drop", None
    if not (raw["score"].max() > model.start_threshold):
        return "drop", None
    t3 = (stop_position > seq_len - 2) if frame > 0 else (stop_position < 1)
    n_len = model.len_lo.shape[3]
    tol = 1e-9
    cand = {}
    for i in range(n):
        r = raw[i]
        if 1 + r["j"] < min_gene_len:
            continue
        length = (1 + int(r["j"])) // 3
        if length >= n_len:
            return "undecided", None
        x = np.float64(r["score"]) + np.float64(model.prior)
        if r["which"] >= 0:
            x = x + np.float64(model.start_lo[int(r["which"])])
        x = x + np.float64(model.len_lo[seq_cls, 1 if r["truncated"] else 0, 1 if t3 else 0, length])
        # candidates that can pass `ne->score > Event_Threshold` once the (>= 0) RBS term is added; the band of width
        # tol below the threshold is kept for the host's exact test
        if x + model.pwm_bonus_max > model.event_threshold - tol:
            cand.setdefault(int(r["pos"]), []).append((x, i))
    win = {}
    for p, lst in cand.items():
        top = max(x for x, _ in lst)
        near = [i for x, i in lst if x >= top - tol]
        if len(near) > 1:  # a tie (or close enough for the RBS term's rounding to decide): std::sort's order matters
            return "undecided", None
        win[p] = near[0]
    return "keep", win


def check_reduction(orfs, ooff, seq_lens, raw_starts, soff, red, first, cnt, status, min_gene_len, model, orf_ids=None):
    """Device reduction (gmg_reduce_starts_mg) against reduce_reference for the ORFs `orf_ids` (default: all).
    Status 2 is always acceptable where the reference outcome is 'undecided'; elsewhere the survivors must be
    exactly the reference's winners (same records, byte for byte)."""
    seq_of = np.repeat(np.arange(len(ooff) - 1), np.diff(ooff))
    stats = {"orfs": 0, "kept_orfs": 0, "dropped_orfs": 0, "handed_back": 0, "survivors": 0, "raw": 0}
    ids = range(len(orfs)) if orf_ids is None else orf_ids
    for o in ids:
        raw = raw_starts[int(soff[o]):int(soff[o + 1])]
        what, win = reduce_reference(raw, int(orfs[o]["frame"]), int(orfs[o]["stop_position"]), int(seq_lens[seq_of[o]]),
                                     min_gene_len, model)
        stats["orfs"] += 1
        stats["raw"] += len(raw)
        st = int(status[o])
        if st == 2:
            stats["handed_back"] += 1
            _require(what == "undecided" or len(raw) > 0, f"ORF {o}: handed back without records")
            continue
        _require(what != "undecided", f"ORF {o}: the outcome depends on sort order / a tie but the device decided it (status {st})")
        if what == "drop" or not win:
            _require(st == 0 or int(cnt[o]) == 0, f"ORF {o}: reference drops it, device status {st} with {int(cnt[o])} survivors")
            stats["dropped_orfs"] += 1
            continue
        _require(st == 1, f"ORF {o}: reference keeps {len(win)} starts, device status {st}")
        got = red[int(first[o]):int(first[o]) + int(cnt[o])]
        _require(len(got) == len(win), f"ORF {o}: {len(got)} survivors, reference {len(win)}")
        bypos = {int(g["pos"]): g for g in got}
        _require(len(bypos) == len(got), f"ORF {o}: two survivors at one position")
        for p, i in win.items():
            _require(p in bypos and bypos[p].tobytes() == raw[i].tobytes(), f"ORF {o}: survivor at position {p} differs")
        stats["kept_orfs"] += 1
        stats["survivors"] += len(got)
    return stats
