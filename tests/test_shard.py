"""Host-side multi-GPU logic on CPU: sharding helpers, and the training exchange step with two gloo ranks.

The device data path needs a B200; what runs here is everything around it: how sequences are dealt to ranks
and that summing per-rank count slabs (the all-reduce the trainer asks for, icm.cc:1092-1093 makes counts
additive over strings) reproduces the single-process counts and therefore the single-process model.  The
counter is the oracle (test infrastructure) standing in for K4."""
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from glimmer_mg_b200 import shard

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_round_robin_partitions():
    for n, w in [(0, 2), (1, 2), (7, 3), (500500, 8)]:
        got = np.concatenate([shard.round_robin(n, r, w) for r in range(w)])
        assert sorted(got.tolist()) == list(range(n))
    with pytest.raises(ValueError):
        shard.round_robin(4, 2, 2)


def test_balanced_ranges_cover_in_order():
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 900, size=1000)
    off = np.zeros(1001, np.int64)
    off[1:] = np.cumsum(lens)
    for w in (1, 2, 4, 8, 1500):
        cut = shard.balanced_ranges(off, w)
        assert cut[0] == 0 and cut[-1] == 1000 and (np.diff(cut) >= 0).all()
        if w <= 8:
            per = np.array([off[cut[r + 1]] - off[cut[r]] for r in range(w)])
            assert per.max() - per.min() <= 2 * lens.max()
    # empty batch and single sequence
    assert shard.balanced_ranges(np.zeros(1, np.int64), 4).tolist() == [0, 0, 0, 0, 0]
    assert shard.balanced_ranges(np.array([0, 10]), 2)[-1] == 1


def test_take_and_slice():
    seqs = [b"acgt", b"", b"ttgacc", b"a"]
    a = np.frombuffer(b"".join(seqs), np.uint8)
    off = np.array([0, 4, 4, 10, 11], np.int64)
    s, o = shard.take_sequences(a, off, [2, 0])
    assert s.tobytes() == b"ttgaccacgt" and o.tolist() == [0, 6, 10]
    s, o = shard.slice_range(a, off, 1, 3)
    assert s.tobytes() == b"ttgacc" and o.tolist() == [0, 0, 6]
    eq = np.frombuffer(b"aaacccgggttt", np.uint8)
    s, o = shard.take_sequences(eq, np.arange(5) * 3, shard.round_robin(4, 1, 2))
    assert s.tobytes() == b"cccttt" and o.tolist() == [0, 3, 6]


def test_group_by_model_is_stable():
    g = shard.group_by_model([2, 0, 2, 1, 0])
    assert {k: v.tolist() for k, v in g.items()} == {0: [1, 4], 1: [3], 2: [0, 2]}


def _rank_main(rank, world, port, q):
    """One gloo rank: count every level on this rank's round-robin shard, all-reduce, compare to the whole."""
    try:
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        strs = [s.lower()[::-1] for _, s in O.read_fasta(os.path.join(G, "seqs.cluster-5.run1.filt.gene.fasta.gz"))]
        mine = [strs[i] for i in shard.round_robin(len(strs), rank, world)]
        arr_all, arr_mine = O.cstr_array(strs), O.cstr_array(mine)
        whole = O.lib().orc_icm_train(arr_all, len(strs), 12, 7, 3)
        omip, _ = O.icm_tables(whole)
        om = O.lib().orc_icm_new(12, 7, 3)
        N = 21845
        for level in range(8):
            nl, first = 4 ** level, (4 ** level - 1) // 3
            local = np.zeros(3 * N * 11 * 16, np.int32)
            O.lib().orc_count_level(om, arr_mine, len(mine), level, local.ctypes.data)
            slab = np.ascontiguousarray(local.reshape(3, N, 11, 16)[:, first:first + nl])
            t = torch.from_numpy(slab)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the exchange step
            full = np.zeros(3 * N * 11 * 16, np.int32)
            O.lib().orc_count_level(om, arr_all, len(strs), level, full.ctypes.data)
            assert (slab == full.reshape(3, N, 11, 16)[:, first:first + nl]).all(), f"level {level}"
            for f in range(3):
                for i in range(first, first + nl):
                    om.contents.mip[f * N + i] = int(omip[f, i])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))


def test_training_exchange_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res


def _score_rank(rank, world, port, q):
    """Scoring shards with no collective: each rank's ORF tables for its base-balanced range, gathered in
    rank order, equal the single-process tables (the oracle stands in for the device path)."""
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        reads = [O.filter_lower(s) for _, s in O.read_fasta(os.path.join(G, "seqs.fa.gz"))[:40]]
        off = np.zeros(len(reads) + 1, np.int64)
        off[1:] = np.cumsum([len(r) for r in reads])
        cut = shard.balanced_ranges(off, world)
        p = O.params(True)
        mine = [O.find_orfs(reads[i], p).tolist() for i in range(cut[rank], cut[rank + 1])]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)  # host-side gather of variable-length results
        if rank == 0:
            flat = [x for part in gathered for x in part]
            want = [O.find_orfs(r, p).tolist() for r in reads]
            assert flat == want
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))


def test_scoring_shards_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_score_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: "ok", 1: "ok"}, res
