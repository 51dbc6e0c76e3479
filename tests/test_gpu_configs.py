"""Parity gates on BASELINE.json configs 2-5 (the synthetic workloads every bench number is quoted on).

Scaled instances of tools/workloads.py's generators with the BASELINE seeds, through the C-ABI, against
  * the oracle port (ORF tables, start lists incl. FP64 score bits, model tables) -- always, and
  * the unmodified reference binaries of oracle/_ref (`.predict` byte compare through the reference's own drivers
    compiled against the C-ABI, model files byte compare) -- where oracle/_ref was built.
BASELINE.md section 3 asks for >= 95 % identical .predict lines; the assertion here is 100 %.
"""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
import config_parity as CP

sys.path.insert(0, os.path.join(O.ROOT, "tools"))
import workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu
ICM_PATH = os.path.join(O.GOLDEN, "NC_000915.icm")
STOPS = ("taa", "tag", "tga")


@pytest.fixture(scope="module")
def gm():
    import glimmer_mg_b200 as g
    return g


@pytest.fixture(scope="module")
def ctx(gm):
    c = gm.Context(0)
    yield c
    c.close()


def _need_ref(*names):
    for n in names:
        if not CP.have_ref_bin(n):
            pytest.skip(f"oracle/_ref/bin/{n} not built (needs the reference checkout at build time)")


# ---------------------------------------------------------------- config 2: contig, glimmer3
def test_config2_contig_matches_oracle(gm, ctx):
    """500 kbp of the config-2 generator (seed 20261017): ORF table and every start list bit-identical."""
    contig = W.contig(W.CONTIG_SEED, 500_000)
    off = np.array([0, len(contig)], np.int64)
    ss = gm.SeqSet(ctx, ascii=contig, offsets=off)
    gc = ss.gc_fraction()
    p = gm.Params(False)
    p.set_ignore_score_len(gc)
    gene = gm.ICM.Read(ctx, ICM_PATH)
    indep = gm.ICM.Build_Indep_WO_Stops(ctx, gc, STOPS)
    ss.find_orfs(p)
    ss.score_orfs_g3(gene, indep, p)
    orfs, ooff = ss.get_orfs()
    starts, soff = ss.get_starts()
    st = CP.check_scoring("g3", contig, off, [0], orfs, ooff, starts, soff, CP.oracle_model(path=ICM_PATH), gc, STOPS,
                          ignore_score_len=p.ignore_score_len)
    assert st["orfs"] > 3000 and st["starts"] > 10000
    assert ss.ordered_fallbacks == 0


def test_config2_contig_predict_equals_reference_binary(tmp_path):
    _need_ref("glimmer3", "glimmer3-gmg")
    contig = W.contig(W.CONTIG_SEED, 500_000)
    fa = str(tmp_path / "contig.fa")
    W.write_fasta(fa, contig, prefix="contig")
    ref, got = CP.predict_pair("glimmer3", ["-u", "-12", "-m", ICM_PATH], fa, str(tmp_path))
    assert ref.count(b"orf") >= 10  # the synthetic genes rarely beat the sample genome's ICM: few calls, all must agree
    assert CP.predict_identity(ref, got) == 1.0 and ref == got


# ---------------------------------------------------------------- config 3: 400 bp reads, -i
@pytest.fixture(scope="module")
def reads400():
    contig = W.contig(W.CONTIG_SEED, 5_000_000)
    return W.reads(contig, 2000, 400, W.READS400_SEED, indel=True)


def test_config3_reads400_indel_matches_oracle(gm, ctx, reads400):
    a, off = reads400
    ss = gm.SeqSet(ctx, ascii=a, offsets=off)
    gc = ss.gc_fraction()
    p = gm.Params(True, allow_indels=1)
    p.set_ignore_score_len(gc)
    gene = gm.ICM.Read(ctx, ICM_PATH)
    indep = gm.ICM.Build_Indep_WO_Stops(ctx, gc, STOPS)
    ss.find_orfs(p)
    og = CP.oracle_model(path=ICM_PATH)
    ss.score_orfs_mg(gene, indep, p)
    orfs, ooff = ss.get_orfs()
    starts, soff = ss.get_starts()
    st = CP.check_scoring("mg", a, off, list(range(len(off) - 1)), orfs, ooff, starts, soff, og, gc, STOPS, allow_indels=1,
                          ignore_score_len=p.ignore_score_len)
    assert st["starts"] > 100 * st["seqs"]
    # row a11b: the device-side reduction of those lists against the reference's filter restated on the raw lists
    model = gm.EventModel(prior=0.0)
    n_kept = ss.reduce_starts_mg(p, model)
    red, first, cnt, status = ss.get_reduced_starts()
    rs = CP.check_reduction(orfs, ooff, np.diff(off), starts, soff, red, first, cnt, status, p.min_gene_len, model,
                            orf_ids=range(0, len(orfs), 7))
    assert rs["kept_orfs"] > 100 and n_kept * 5 < len(starts), (rs, n_kept, len(starts))
    assert ss.uncertified == 0


def test_config3_reads400_predict_equals_reference_binary(tmp_path, reads400):
    _need_ref("glimmer-mg", "glimmer-mg-gmg")
    a, off = reads400
    n = 600
    fa = str(tmp_path / "reads.fa")
    W.write_fasta(fa, a[:off[n]], off[:n + 1], prefix="r")
    ref, got = CP.predict_pair("glimmer-mg", ["-u", "1.0", "-i", "-m", ICM_PATH], fa, str(tmp_path))
    assert ref.count(b"orf") > 300 and b" I:" in ref
    assert CP.predict_identity(ref, got) == 1.0 and ref == got


# ---------------------------------------------------------------- config 5: 100 bp reads, per-cluster ICMs
def _cluster(k, n_clusters=16):
    gc = float(np.linspace(0.30, 0.70, n_clusters)[k])
    freq = W.reweight_gc(W.codon_freq(), gc)
    genome = W.contig(W.READS100_SEED * 1000 + k, 400_000, freq=freq, gc=gc)
    train = W.coding(1500, 333, seed=W.READS100_SEED * 1000 + 500 + k, freq=freq)
    return genome, train


@pytest.mark.parametrize("k", [0, 15])
def test_config5_reads100_cluster_icm_matches_oracle(gm, ctx, tmp_path, k):
    """One cluster of config 5: the ICM trained on the device equals the oracle's (and the reference build-icm's) byte
    for byte, and 10 000 error-free 100 bp reads scored with it give the oracle's ORFs and start lists."""
    genome, (ts, toff) = _cluster(k)
    model = gm.ICMTraining(ctx, 12, 7, 3).Train_Model(gm.SeqSet(ctx, ascii=ts, offsets=toff), reverse=True)
    mpath = str(tmp_path / f"cluster{k}.icm")
    model.Output(mpath)
    raw = ts.tobytes()
    rev = [raw[toff[i]:toff[i + 1]][::-1] for i in range(len(toff) - 1)]
    om = O.lib().orc_icm_train(O.cstr_array(rev), len(rev), 12, 7, 3)
    opath = str(tmp_path / f"oracle{k}.icm")
    O.lib().orc_icm_write(om, opath.encode())
    assert open(mpath, "rb").read() == open(opath, "rb").read(), "device-trained cluster ICM differs from the oracle's"
    if CP.have_ref_bin("build-icm"):
        tfa = str(tmp_path / "train.fa")
        W.write_fasta(tfa, ts, toff, prefix="g")
        rpath = str(tmp_path / f"ref{k}.icm")
        CP.run([os.path.join(CP.REFBIN, "build-icm"), "-r", rpath], stdin_path=tfa)
        assert open(mpath, "rb").read() == open(rpath, "rb").read(), "cluster ICM differs from the reference build-icm's"
    a, off = W.reads(genome, 10_000, 100, W.READS100_SEED + 7919 * k, indel=False)
    ss = gm.SeqSet(ctx, ascii=a, offsets=off)
    gc = ss.gc_fraction()
    p = gm.Params(True)
    p.set_ignore_score_len(gc)
    indep = gm.ICM.Build_Indep_WO_Stops(ctx, gc, STOPS)
    ss.find_orfs(p)
    ss.score_orfs_mg(model, indep, p)
    orfs, ooff = ss.get_orfs()
    starts, soff = ss.get_starts()
    st = CP.check_scoring("mg", a, off, range(len(off) - 1), orfs, ooff, starts, soff, om, gc, STOPS,
                          ignore_score_len=p.ignore_score_len)
    assert st["orfs"] > 1000 and ss.uncertified == 0
    if CP.have_ref_bin("glimmer-mg") and CP.have_ref_bin("glimmer-mg-gmg"):
        fa = str(tmp_path / "reads.fa")
        W.write_fasta(fa, a[:off[3000]], off[:3001], prefix="r")
        ref, got = CP.predict_pair("glimmer-mg", ["-u", "1.0", "-m", mpath], fa, str(tmp_path))
        assert CP.predict_identity(ref, got) == 1.0 and ref == got


# ---------------------------------------------------------------- config 4: training
@pytest.mark.parametrize("hist", ["0", "1"])
def test_config4_training_model_file_equals_reference(gm, ctx, tmp_path, monkeypatch, hist):
    """5 Mbp of config 4 (seed 7) through both counting paths (direct / window histogram): model file byte-identical
    to the reference build-icm binary's (oracle port when oracle/_ref is absent)."""
    monkeypatch.setenv("GMG_K4_HIST", hist)
    ts, toff = W.coding(5005, 333, W.TRAIN_SEED)
    model = gm.ICMTraining(ctx, 12, 7, 3).Train_Model(gm.SeqSet(ctx, ascii=ts, offsets=toff), reverse=True)
    mpath = str(tmp_path / "dev.icm")
    model.Output(mpath)
    want = str(tmp_path / "want.icm")
    if CP.have_ref_bin("build-icm"):
        tfa = str(tmp_path / "train.fa")
        W.write_fasta(tfa, ts, toff, prefix="g")
        CP.run([os.path.join(CP.REFBIN, "build-icm"), "-r", want], stdin_path=tfa)
    else:
        raw = ts.tobytes()
        rev = [raw[toff[i]:toff[i + 1]][::-1] for i in range(len(toff) - 1)]
        O.lib().orc_icm_write(O.lib().orc_icm_train(O.cstr_array(rev), len(rev), 12, 7, 3), want.encode())
    assert CP.file_sha256(mpath) == CP.file_sha256(want)
