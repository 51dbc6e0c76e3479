"""Parser for the raw ORF / start-list dumps written by the instrumented reference
drivers (oracle/mg_dump_hook.inc) -- see tests/golden/make_golden.py."""
import gzip

import numpy as np


def parse_dump(path):
    """-> list of reads: dict(len, hdr, fs=[6 x uint64 arrays] or [], orfs=[dict(o=(frame, stop, orf_len,
    gene_len), starts=[(j, pos, score_bits, which, truncated, first, errors)])])."""
    op = gzip.open if str(path).endswith(".gz") else open
    recs, cur = [], None
    with op(path, "rt") as f:
        for line in f:
            t = line.split()
            if t[0] == "R":
                cur = dict(len=int(t[1]), hdr=t[2], fs=[], orfs=[])
                recs.append(cur)
            elif t[0] == "F":
                cur["fs"].append(np.array([int(x, 16) for x in t[2:]], np.uint64))
            elif t[0] == "O":
                cur["orfs"].append(dict(o=tuple(int(x) for x in t[1:5]), starts=[]))
            elif t[0] == "S":
                errs = tuple(tuple(int(y) for y in x.split(":")) for x in t[8:])
                cur["orfs"][-1]["starts"].append(
                    (int(t[1]), int(t[2]), int(t[3], 16), int(t[4]), int(t[5]), int(t[6]), errs))
    return recs


def boost(starts, ignore_score_len):
    """long-ORF boost (glimmer-mg.cc:1649-1651) applied to dumped (pre-boost) starts."""
    out = []
    for (j, pos, sc, w, tr, fi, er) in starts:
        if j > ignore_score_len and 0.0 > np.uint64(sc).view(np.float64):
            sc = 0
        out.append((j, pos, sc, w, tr, fi, er))
    return out


def starts_as_tuples(st):
    """START_DTYPE-like structured array (fields j,pos,score,which,truncated,first,n_err,err_pos,err_type)."""
    return [(int(x["j"]), int(x["pos"]), int(np.float64(x["score"]).view(np.uint64)), int(x["which"]),
             int(x["truncated"]), int(x["first"]),
             tuple((int(x["err_pos"][e]), int(x["err_type"][e])) for e in range(int(x["n_err"])))) for x in st]
