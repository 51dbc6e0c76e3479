"""glimmer_mg_b200/host/glimmer-mg.py keeps the option surface of the reference's pipeline driver
(scripts/glimmer-mg.py:142-193) and builds the reference's glimmer-mg / build-icm command lines."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "glimmer_mg_b200", "host", "glimmer-mg.py")
REF = "/root/reference/scripts/glimmer-mg.py"

# the reference's option strings (scripts/glimmer-mg.py:146-192), committed so that the test runs without the checkout
REF_OPTIONS = {"--iter", "--long_orfs", "-o", "-p", "--single_cluster", "-t", "--filter", "--glim_bin", "--ignore", "--all_features",
               "--time", "--skip_first", "-i", "--indel", "-q", "-r", "--circular", "-s", "--sub", "-u", "--fudge", "--raw", "--class",
               "--clust", "--taxlevel", "--minbp_pct"}


def _mod():
    spec = importlib.util.spec_from_file_location("gmg_frontend", PATH)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_option_surface_matches_the_reference():
    ap = _mod().build_parser()
    mine = {s for a in ap._actions for s in a.option_strings}
    assert REF_OPTIONS <= mine, REF_OPTIONS - mine
    if os.path.exists(REF):  # the committed list is the reference's
        src = "".join(l for l in open(REF) if not l.lstrip().startswith("#"))  # '-g' is commented out there
        found = set(re.findall(r"add_option\(\s*'(-{1,2}[A-Za-z_]+)'(?:\s*,\s*'(--[A-Za-z_]+)')?", src))
        ref = {x for pair in found for x in pair if x}
        assert ref == REF_OPTIONS, ref ^ REF_OPTIONS
    d = {a.dest: a.default for a in ap._actions}
    assert (d["iterate"], d["proc"], d["top_hits"], d["fudge"], d["taxlevel"], d["minbp_pct"], d["filter_t"]) == (1, 1, 3, 1.0, "family",
                                                                                                                  0.01, 1.0)


def test_command_lines(capsys):
    m = _mod()
    log = m.main(["--dry_run", "--class", "--iter", "0", "-i", "-s", "-u", "0.5", "-q", "r.qual", "-o", "out", "reads.fa"])
    assert log == [f"{os.path.join(ROOT, 'glimmer_mg_b200', 'host', 'bin', 'glimmer-mg-gmg')} -u 0.500000 -i -s -c out.class.txt -q r.qual reads.fa out"]
    log = m.main(["--dry_run", "--class", "--single_cluster", "--iter", "1", "--long_orfs", "reads.fa"])
    assert log[2].endswith("build-icm -r reads.run1.icm < reads.train")
    assert log[3].endswith("-u 1.000000 -m reads.run1.icm -c reads.class.txt reads.fa reads.run1")
    assert log[-1].endswith("-b reads.run1.motif -m reads.run1.gicm -f reads.run1.features.txt -c reads.class.txt reads.fa reads")
    with pytest.raises(SystemExit, match="linear"):
        m.main(["--dry_run", "--class", "-r", "reads.fa"])
