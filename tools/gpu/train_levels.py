"""Per-level breakdown of Train_Model on the train500m workload: K4 device time vs host finish_level time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import glimmer_mg_b200 as g
import workloads as W
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = g.Context(0, stream.cuda_stream)
    a, off = W.coding(int(500500 * scale), 333, W.TRAIN_SEED)
    ss = g.SeqSet(ctx, ascii=a, offsets=off)
    for rep in range(2):
        tr = g.ICMTraining(ctx, 12, 7, 3).levels(ss, reverse=True)
        for level in range(8):
            ctx.sync(); t0 = time.perf_counter()
            tr.count_level(level)
            ctx.sync(); t1 = time.perf_counter()
            tr.finish_level(level)
            ctx.sync(); t2 = time.perf_counter()
            if rep: print(f"level {level}: count {1e3*(t1-t0):8.2f} ms   finish {1e3*(t2-t1):8.2f} ms")
        t0 = time.perf_counter(); m = tr.finish(); ctx.sync(); print(f"finish: {1e3*(time.perf_counter()-t0):.2f} ms")
        tr.close()
