"""Where the end-to-end time of bench.py's reads100 / reads400 step goes: every C-ABI call timed with a stream sync."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import glimmer_mg_b200 as g, workloads as W
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "reads100"
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = g.Context(0, stream.cuda_stream)
    batches, desc = bench.reads_batches(kind, 0, 1, 1.0, 1)
    a, off, k = batches[0]
    h = torch.empty(len(a), dtype=torch.uint8).pin_memory(); h.numpy()[:] = a
    gene = g.ICM.Read(ctx, W.gene_model_path())
    p = g.Params(True, allow_indels=1 if kind == "reads400" else 0); indep = None
    acc = {}
    def T(name, fn):
        ctx.sync(); t = time.perf_counter(); r = fn(); ctx.sync(); acc.setdefault(name, []).append(time.perf_counter() - t); return r
    for it in range(8):
        ss = T("seqset_create", lambda: g.SeqSet(ctx, ascii=h.numpy(), offsets=off))
        gc = T("gc_fraction", ss.gc_fraction)
        if indep is None:
            p.set_ignore_score_len(gc); indep = g.ICM.Build_Indep_WO_Stops(ctx, gc, p.stop_codons)
        T("find_orfs", lambda: ss.find_orfs(p))
        T("score_orfs_mg", lambda: ss.score_orfs_mg(gene, indep, p))
        T("get_orfs", lambda: ss.get_orfs(pinned=True))
        em = g.EventModel(prior=0.0, len_lo=np.zeros((1, 2, 2, 256)))
        T("reduce_starts_mg", lambda: ss.reduce_starts_mg(p, em))
        T("get_reduced_starts", lambda: ss.get_reduced_starts(pinned=True))
        T("close", ss.close)
    for k2, v in acc.items():
        print(f"{k2:22s} first {v[0]*1e3:8.3f} ms   steady {np.median(v[3:])*1e3:8.3f} ms")
    print("n_orfs", ss.n_orfs, "n_starts", ss.n_starts, "bases", int(off[-1]))
