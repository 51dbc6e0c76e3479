"""CPU-side checks of the drop-in boundary: libgmgicm.so loads without a GPU and exports every symbol that
include/gmg_icm.h declares; the Python mirror declares the same set; no compute call is made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gmg_icm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gmg_[a-z0-9_]+)\s*\(", src))
    names.discard("gmg_allreduce_fn")
    return names


def test_header_symbols_are_exported():
    import glimmer_mg_b200 as g
    lib = C.CDLL(g.lib_path())
    names = _declared()
    assert len(names) >= 40
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/gmg_icm.h but not exported by libgmgicm.so"


def test_python_mirror_binds_every_symbol():
    import glimmer_mg_b200 as g
    L = g.lib()  # sets argtypes/restype for every entry; AttributeError if one is missing
    assert L.gmg_abi_version() == 1
    src = open(os.path.join(ROOT, "glimmer_mg_b200", "icm.py")).read()
    bound = set(re.findall(r'"(gmg_[a-z0-9_]+)"\s*:', src))
    assert _declared() <= bound, sorted(_declared() - bound)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to create a context (and says so) instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import glimmer_mg_b200 as g
    with pytest.raises(g.GmgError, match="no usable CUDA device|no CPU fallback"):
        g.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under glimmer_mg_b200/ may import, link or call oracle/ (the oracle is the checker only)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "glimmer_mg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)


def test_params_defaults():
    import glimmer_mg_b200 as g
    p = g.Params(True)
    assert (p.min_gene_len, p.allow_truncated, p.min_indel_orf_len, p.indel_quality_threshold, p.indel_max) == (75, 1, 15, 18, 2)
    assert p.stop_codons == ["taa", "tag", "tga"]
    assert g.Params(False).allow_truncated == 0
    # Set_Ignore_Score_Len (glimmer_base.cc:2597-2633) at the sample genome's GC
    assert p.set_ignore_score_len(0.3887516210824964) > 0
