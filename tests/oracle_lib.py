"""ctypes bindings for the CHECKERS (test infrastructure only):

* ``oracle/liboracle.so``  -- our plain-C restatement (oracle/icm_oracle.c)
* ``oracle/_ref/lib/libref_icm.so`` -- the unmodified reference ICM library behind
  oracle/ref_shim.cc (present only where oracle/_ref was built)

Nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_SAMPLE = "/root/reference/sample-run"
GOLDEN = os.path.join(ROOT, "tests", "golden")


class OrcIcm(C.Structure):
    _fields_ = [("model_len", C.c_int), ("model_depth", C.c_int), ("periodicity", C.c_int),
                ("num_nodes", C.c_int), ("mip", C.POINTER(C.c_short)), ("prob", C.POINTER(C.c_float))]


class OrcOrf(C.Structure):
    _fields_ = [("frame", C.c_int), ("stop_position", C.c_int), ("orf_len", C.c_int), ("gene_len", C.c_int)]


class OrcStart(C.Structure):
    _fields_ = [("j", C.c_int), ("pos", C.c_int), ("score", C.c_double), ("which", C.c_int),
                ("truncated", C.c_int), ("first", C.c_int), ("n_err", C.c_int),
                ("err_pos", C.c_int * 2), ("err_type", C.c_int * 2)]


class OrcParams(C.Structure):
    _fields_ = [("min_gene_len", C.c_int), ("allow_truncated", C.c_int), ("allow_indels", C.c_int),
                ("allow_subs", C.c_int), ("min_indel_orf_len", C.c_int), ("indel_quality_threshold", C.c_int),
                ("indel_max", C.c_int), ("indel_suffix_score_threshold", C.c_double),
                ("ignore_score_len", C.c_int), ("have_quality_file", C.c_int), ("n_start", C.c_int),
                ("n_stop", C.c_int), ("start_codon", (C.c_char * 4) * 8), ("stop_codon", (C.c_char * 4) * 8)]


ORF_DTYPE = np.dtype([("frame", "<i4"), ("stop_position", "<i4"), ("orf_len", "<i4"), ("gene_len", "<i4")])
START_DTYPE = np.dtype([("j", "<i4"), ("pos", "<i4"), ("score", "<f8"), ("which", "<i4"), ("truncated", "<i4"),
                        ("first", "<i4"), ("n_err", "<i4"), ("err_pos", "<i4", (2,)), ("err_type", "<i4", (2,))])
assert START_DTYPE.itemsize == C.sizeof(OrcStart)

_lib = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "icm_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build_oracle()
    L = C.CDLL(path)
    P = C.POINTER
    L.orc_default_params.argtypes = [P(OrcParams), C.c_int]
    L.orc_icm_new.restype = P(OrcIcm)
    L.orc_icm_new.argtypes = [C.c_int] * 3
    L.orc_icm_read.restype = P(OrcIcm)
    L.orc_icm_read.argtypes = [C.c_char_p]
    L.orc_icm_write.argtypes = [P(OrcIcm), C.c_char_p]
    L.orc_icm_free.argtypes = [P(OrcIcm)]
    L.orc_build_indep_wo_stops.restype = P(OrcIcm)
    L.orc_build_indep_wo_stops.argtypes = [C.c_double, P(C.c_char_p), C.c_int]
    L.orc_full_window_prob.restype = C.c_double
    L.orc_full_window_prob.argtypes = [P(OrcIcm), C.c_char_p, C.c_int]
    L.orc_partial_window_prob.restype = C.c_double
    L.orc_partial_window_prob.argtypes = [P(OrcIcm), C.c_int, C.c_char_p, C.c_int]
    L.orc_score_string.restype = C.c_double
    L.orc_score_string.argtypes = [P(OrcIcm), C.c_char_p, C.c_int, C.c_int]
    L.orc_cumulative_score.argtypes = [P(OrcIcm), C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_frame_score.argtypes = [P(OrcIcm), C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    L.orc_score_all_frames.argtypes = [P(OrcIcm), P(OrcIcm), C.c_char_p, C.c_int, C.c_void_p]
    L.orc_save_prev_stops.argtypes = [C.c_char_p, C.c_int, P(OrcParams), C.c_void_p, C.c_void_p]
    L.orc_set_quality_454.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    L.orc_gc_fraction.restype = C.c_double
    L.orc_gc_fraction.argtypes = [P(C.c_char_p), P(C.c_int), C.c_int]
    L.orc_ignore_score_len.argtypes = [C.c_double, P(OrcParams)]
    L.orc_find_orfs.argtypes = [C.c_char_p, C.c_int, P(OrcParams), P(P(OrcOrf))]
    L.orc_mg_score_orfs.argtypes = [P(OrcIcm), P(OrcIcm), C.c_char_p, C.c_int, C.c_void_p, P(OrcParams),
                                    C.c_void_p, C.c_int, C.c_void_p, P(P(OrcStart))]
    L.orc_g3_score_orfs.argtypes = [P(OrcIcm), P(OrcIcm), C.c_char_p, C.c_int, P(OrcParams),
                                    C.c_void_p, C.c_int, C.c_void_p, P(P(OrcStart))]
    L.orc_icm_train.restype = P(OrcIcm)
    L.orc_icm_train.argtypes = [P(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_count_level.argtypes = [P(OrcIcm), P(C.c_char_p), C.c_int, C.c_int, C.c_void_p]
    L.free = C.CDLL(None).free
    L.free.argtypes = [C.c_void_p]
    _lib = L
    return L


def params(metagenomic=True, **kw):
    p = OrcParams()
    lib().orc_default_params(C.byref(p), 1 if metagenomic else 0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def icm_tables(m):
    """(mip int16 [P,N], prob float32 [P,N,4]) copies of an orc_icm*."""
    c = m.contents
    n = c.periodicity * c.num_nodes
    mip = np.ctypeslib.as_array(c.mip, shape=(n,)).reshape(c.periodicity, c.num_nodes).copy()
    prob = np.ctypeslib.as_array(c.prob, shape=(n * 4,)).reshape(c.periodicity, c.num_nodes, 4).copy()
    return mip, prob


def cstr_array(strings):
    arr = (C.c_char_p * len(strings))()
    for i, s in enumerate(strings):
        arr[i] = s if isinstance(s, bytes) else s.encode()
    return arr


def build_indep(gc, stops=("taa", "tag", "tga")):
    return lib().orc_build_indep_wo_stops(gc, cstr_array(list(stops)), len(stops))


def find_orfs(seq, p):
    out = C.POINTER(OrcOrf)()
    n = lib().orc_find_orfs(seq, len(seq), C.byref(p), C.byref(out))
    if n == 0:
        return np.zeros(0, ORF_DTYPE)
    arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int)), shape=(n * 4,)).copy().view(ORF_DTYPE)
    lib().free(out)
    return arr


def _score_orfs(fn, gene, indep, seq, p, orfs, qual=None):
    n = len(orfs)
    orfs = np.ascontiguousarray(orfs)
    off = np.zeros(n + 1, np.int32)
    out = C.POINTER(OrcStart)()
    if fn == "mg":
        q = None if qual is None else np.ascontiguousarray(qual, np.int32)
        total = lib().orc_mg_score_orfs(gene, indep, seq, len(seq), None if q is None else q.ctypes.data,
                                        C.byref(p), orfs.ctypes.data, n, off.ctypes.data, C.byref(out))
    else:
        total = lib().orc_g3_score_orfs(gene, indep, seq, len(seq), C.byref(p), orfs.ctypes.data, n,
                                        off.ctypes.data, C.byref(out))
    if total == 0:
        starts = np.zeros(0, START_DTYPE)
    else:
        buf = C.string_at(out, total * C.sizeof(OrcStart))
        starts = np.frombuffer(buf, START_DTYPE).copy()
    if out:
        lib().free(out)
    return off, starts


def mg_score_orfs(gene, indep, seq, p, orfs, qual=None):
    return _score_orfs("mg", gene, indep, seq, p, orfs, qual)


def g3_score_orfs(gene, indep, seq, p, orfs):
    return _score_orfs("g3", gene, indep, seq, p, orfs)


def score_all_frames(gene, indep, seq):
    fs = np.zeros((6, len(seq)), np.float64)
    lib().orc_score_all_frames(gene, indep, seq, len(seq), fs.ctypes.data)
    return fs


# --------------------------------------------------------------------------- reference
_ref = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "lib", "libref_icm.so"))


def ref():
    global _ref
    if _ref is not None:
        return _ref
    R = C.CDLL(os.path.join(REF_DIR, "lib", "libref_icm.so"))
    R.ref_icm_read.restype = C.c_void_p
    R.ref_icm_read.argtypes = [C.c_char_p]
    R.ref_icm_build_indep.restype = C.c_void_p
    R.ref_icm_build_indep.argtypes = [C.c_double, C.POINTER(C.c_char_p), C.c_int]
    R.ref_icm_free.argtypes = [C.c_void_p]
    R.ref_icm_dims.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_icm_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_full_window_prob.restype = C.c_double
    R.ref_full_window_prob.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    R.ref_partial_window_prob.restype = C.c_double
    R.ref_partial_window_prob.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    R.ref_score_string.restype = C.c_double
    R.ref_score_string.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    R.ref_cumulative_score.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    R.ref_frame_score.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    R.ref_icm_train.restype = C.c_void_p
    R.ref_icm_train.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_int]
    R.ref_icm_train_free.argtypes = [C.c_void_p]
    R.ref_icm_write.argtypes = [C.c_void_p, C.c_char_p]
    if hasattr(R, "ref_fasta_read_file"):
        R.ref_fasta_read_file.restype = C.c_int
        R.ref_fasta_read_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_long), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_long)]
    if hasattr(R, "ref_fasta_qual_read_file"):
        R.ref_fasta_qual_read_file.restype = C.c_int
        R.ref_fasta_qual_read_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_long), C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_long)]
    _ref = R
    return R


def ref_qual_records(path):
    """[(header bytes, [values])] as the UNMODIFIED reference reader sees a quality file (Fasta_Qual_Vec_Read,
    Common/fasta.cc:115)."""
    R = ref()
    vp, cp, hp, nv, hn = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_long(), C.c_long()
    n = R.ref_fasta_qual_read_file(os.fsencode(path), C.byref(vp), C.byref(nv), C.byref(cp), C.byref(hp), C.byref(hn))
    assert n >= 0, path
    vals = np.frombuffer(C.string_at(vp, nv.value * 4), np.int32).tolist() if nv.value else []
    cnt = np.frombuffer(C.string_at(cp, n * 8), np.int64).tolist() if n else []
    hdrs = C.string_at(hp, hn.value).split(b"\0")[:n]
    free = C.CDLL(None).free
    free.argtypes = [C.c_void_p]
    for ptr in (vp, cp, hp):
        free(ptr)
    out, k = [], 0
    for h, c in zip(hdrs, cnt):
        out.append((h, vals[k:k + c]))
        k += c
    return out


def ref_fasta_records(path):
    """[(header bytes, sequence bytes)] as the UNMODIFIED reference reader sees a file (Fasta_Read, Common/fasta.cc:236)."""
    R = ref()
    sp, hp, sn, hn = C.c_void_p(), C.c_void_p(), C.c_long(), C.c_long()
    n = R.ref_fasta_read_file(os.fsencode(path), C.byref(sp), C.byref(sn), C.byref(hp), C.byref(hn))
    assert n >= 0, path
    seqs = C.string_at(sp, sn.value).split(b"\0")[:n]
    hdrs = C.string_at(hp, hn.value).split(b"\0")[:n]
    free = C.CDLL(None).free
    free.argtypes = [C.c_void_p]
    free(sp)
    free(hp)
    return list(zip(hdrs, seqs))


def ref_tables(h):
    dims = np.zeros(4, np.int32)
    ref().ref_icm_dims(h, dims.ctypes.data)
    _, _, p, n = [int(x) for x in dims]
    mip = np.zeros((p, n), np.int16)
    prob = np.zeros((p, n, 4), np.float32)
    ref().ref_icm_tables(h, mip.ctypes.data, prob.ctypes.data)
    return mip, prob


# --------------------------------------------------------------------------- data helpers
def read_fasta(path):
    """[(header, sequence-bytes)] -- plain multi-FASTA (gz ok)."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    recs, hdr, parts = [], None, []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip()
            if line.startswith(b">"):
                if hdr is not None:
                    recs.append((hdr, b"".join(parts)))
                hdr, parts = line[1:].decode(), []
            elif hdr is not None:
                parts.append(line)
    if hdr is not None:
        recs.append((hdr, b"".join(parts)))
    return recs


_FILTER = np.full(256, ord("c"), np.uint8)
for _k, _v in dict(a="a", c="c", g="g", t="t", r="g", y="c", s="c", w="t", m="c", k="t", b="c", d="g", h="c",
                   v="c").items():
    _FILTER[ord(_k)] = ord(_v)
    _FILTER[ord(_k.upper())] = ord(_v)


def filter_lower(seq):
    """tolower(Filter(c)) for every character (glimmer-mg.cc:381-382)."""
    return _FILTER[np.frombuffer(seq, np.uint8)].tobytes()


def golden_path(name):
    """A golden file: the committed fixture, else the reference sample-run copy."""
    for cand in (os.path.join(GOLDEN, name), os.path.join(GOLDEN, name + ".gz")):
        if os.path.exists(cand):
            return cand
    for sub in ("glimmer3", "glimmer3/results", "glimmer-mg", "glimmer-mg/results"):
        cand = os.path.join(REF_SAMPLE, sub, name)
        if os.path.exists(cand):
            return cand
    return None


def py_qual_records(image):
    """Fasta_Qual_Vec_Read (Common/fasta.cc:115-170) restated: [(header, [values])]."""
    out, i, n = [], 0, len(image)
    while True:
        while i < n and image[i:i + 1] != b">":
            i += 1
        if i >= n:
            return out
        i += 1
        while i < n and image[i:i + 1] == b" ":
            i += 1
        if i >= n:
            return out
        j = image.find(b"\n", i)
        j = n if j < 0 else j
        hdr, i = image[i:j], min(j + 1, n)
        vals, have, val = [], False, 0
        while i < n and image[i:i + 1] != b">":
            ch = image[i:i + 1]
            if ch.isspace():
                if have:
                    vals.append(val)
                have, val = False, 0
            elif ch.isdigit():
                have, val = True, 10 * val + int(ch)
            i += 1
        out.append((hdr, vals))


QUAL_EDGE_IMAGES = {
    "edge": (b"12 13 junk\n>r1  first \r\n40 40\t7\n\n 0 -3 1x2 999999 0012\n5>r2 mid-line start > odd\n1 2 3\n4"
             b">r3\n>r4\n 8 \x0b9\x0c10\r11 \n>last header only"),
    "edge2": b">   blanks\n1\n2 \n>\n3 4\n> \n\n>x\r5 6\rstill header\n7 8 9\n>tail\n10 11\n>   ",
    "empty": b"",
    "norecord": b"1 2 3\nno records\n",
}
