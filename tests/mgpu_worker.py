"""torchrun worker for tests/test_gpu_multi.py: one process per GPU, NCCL.

(1) Training: every rank trains on its round-robin shard with the per-level count-slab all-reduce; the model
    must equal, byte for byte, the model rank 0 trains alone on all strings (and the oracle's).
(2) Scoring: ranks score base-balanced contiguous ranges of a read set with no collective; the gathered ORF
    tables and start lists must equal rank 0's single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402  (checker)
import glimmer_mg_b200 as g  # noqa: E402
from glimmer_mg_b200 import shard  # noqa: E402

G = os.path.join(HERE, "golden")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = g.Context(local, stream.cuda_stream)
        # ---- (1) training ----
        strs = [s.lower() for _, s in O.read_fasta(os.path.join(G, "NC_000915.train.gz"))]
        mine = [strs[i] for i in shard.round_robin(len(strs), rank, world)]
        total = sum(len(s) for s in strs)
        m = g.ICMTraining(ctx, 12, 7, 3).Train_Model(mine, reverse=True, allreduce=shard.torch_allreduce(local), rank=rank,
                                                     world=world, global_bases=total)
        mip, prob = m.tables()
        blob = mip.tobytes() + prob.tobytes()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        assert all(b == blobs[0] for b in blobs), "ranks hold different models after the all-reduce"
        # the large-set path: window histogram summed over the ranks once, every rank walks 1 / world of its cells
        os.environ["GMG_K4_HIST"] = "1"
        mh = g.ICMTraining(ctx, 12, 7, 3).Train_Model(mine, reverse=True, allreduce=shard.torch_allreduce(local), rank=rank,
                                                      world=world, global_bases=total)
        del os.environ["GMG_K4_HIST"]
        hmip, hprob = mh.tables()
        assert hmip.tobytes() + hprob.tobytes() == blob, "cell-sharded histogram training differs from direct counting"
        if rank == 0:
            alone = g.ICMTraining(ctx, 12, 7, 3).Train_Model(strs, reverse=True)
            amip, aprob = alone.tables()
            assert amip.tobytes() + aprob.tobytes() == blob, "sharded training differs from single-GPU training"
            rev = [s[::-1] for s in strs]
            omip, oprob = O.icm_tables(O.lib().orc_icm_train(O.cstr_array(rev), len(rev), 12, 7, 3))
            assert (omip == mip).all() and (oprob.view(np.uint32) == prob.view(np.uint32)).all(), "differs from the oracle"
        # ---- (2) scoring ----
        reads = [s for _, s in O.read_fasta(os.path.join(G, "seqs.fa.gz"))[:200]]
        off = np.zeros(len(reads) + 1, np.int64)
        off[1:] = np.cumsum([len(r) for r in reads])
        cut = shard.balanced_ranges(off, world)
        gene = g.ICM.Read(ctx, os.path.join(G, "NC_000915.icm"))

        def score(rs):
            ss = g.SeqSet(ctx, seqs=rs)
            p = g.Params(True, allow_indels=1)
            p.set_ignore_score_len(0.39)
            indep = g.ICM.Build_Indep_WO_Stops(ctx, 0.39)
            ss.find_orfs(p)
            ss.score_orfs_mg(gene, indep, p)
            orfs, ooff = ss.get_orfs()
            starts, soff = ss.get_starts()
            return orfs.tobytes(), np.diff(ooff).tolist(), starts.tobytes(), np.diff(soff).tolist()

        part = score(reads[cut[rank]:cut[rank + 1]])
        parts = [None] * world
        dist.all_gather_object(parts, part)
        if rank == 0:
            whole = score(reads)
            assert b"".join(p[0] for p in parts) == whole[0], "ORF tables differ"
            assert sum((p[1] for p in parts), []) == whole[1]
            assert b"".join(p[2] for p in parts) == whole[2], "start lists differ"
            assert sum((p[3] for p in parts), []) == whole[3]
            print(f"mgpu ok: world={world}, model identical on all ranks / single GPU / oracle; "
                  f"{len(whole[1])} reads, {len(whole[3])} ORFs identical after sharding")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
