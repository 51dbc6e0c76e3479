"""The product path has no CPU fallback: without a CUDA device the C-ABI refuses to create a context, and so do the
C++ hosts built on it (our build-icm, and the reference drivers compiled against host/icm.hh) -- they print the
library's message and exit non-zero instead of computing anything on the host."""
import gzip
import os
import subprocess

import pytest

import glimmer_mg_b200 as g

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


pytestmark = pytest.mark.skipif(not _no_gpu(), reason="a GPU is present")


def test_context_creation_fails_loudly():
    with pytest.raises(g.GmgError) as e:
        g.Context(0)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_build_icm_host_refuses_without_gpu(tmp_path):
    exe = os.path.join(ROOT, "glimmer_mg_b200", "host", "bin", "build-icm")
    if not os.path.exists(exe):
        pytest.skip("host/bin/build-icm not built")
    r = subprocess.run([exe, "-r", str(tmp_path / "m.icm")], input=b">a\nacgtacgtacgtacgtacgtacgt\n", capture_output=True)
    assert r.returncode != 0 and b"ERROR" in r.stderr
    assert not os.path.exists(tmp_path / "m.icm") or os.path.getsize(tmp_path / "m.icm") == 0


@pytest.mark.parametrize("name", ["glimmer3-gmg", "glimmer-mg-gmg"])
def test_dropin_drivers_link_the_library_and_refuse_without_gpu(tmp_path, name):
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", name)
    if not os.path.exists(exe):
        pytest.skip(f"{name} not built (needs the reference checkout at build time)")
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libgmgicm.so" in ldd and "not found" not in ldd
    gen = os.path.join(ROOT, "oracle", "_ref", "gen", name.replace("-gmg", "-gmg.cc"))
    src = open(gen, errors="replace").read()
    hook = "Gmg_Score_Orfs (orf_list, gene_list, detail_fp);" if name == "glimmer3-gmg" else \
        "Gmg_Score_Orfs_Errors (orf_list, detail_fp);"
    assert hook in src  # main() calls the binding, not the reference's CPU scoring loop
    fa = tmp_path / "in.fa"
    lines = gzip.open(os.path.join(G, "NC_000915.fna.gz"), "rt").readlines()
    fa.write_text(lines[0] + "".join(lines[1:200]))
    flags = ["-u", "-12"] if name == "glimmer3-gmg" else ["-u", "1.0"]
    r = subprocess.run([exe, *flags, "-m", os.path.join(G, "NC_000915.icm"), str(fa), str(tmp_path / "out")],
                       capture_output=True, timeout=120)
    assert r.returncode != 0 and b"ERROR" in r.stderr
