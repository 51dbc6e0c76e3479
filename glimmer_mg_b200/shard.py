"""Host-side sharding of the hot path across the GPUs of one box (one process per GPU).

Scoring: reads / contigs never interact (glimmer-mg.cc:367-449 handles one read at a time), so a rank takes
a contiguous, base-balanced range of the job's sequences and no collective is needed; per-rank results are
concatenated in input order.  Reads are grouped by ICM first because the reference is ICM-major
(glimmer-mg.cc:361-367; scripts/glimmer-mg.py:127-131 runs one glimmer-mg per cluster).

Training: counts are additive over training strings (icm.cc:1092-1093), so strings are dealt round-robin and
each tree level's int32 count slab is summed across ranks -- the path's one exchange step -- before every rank
runs the identical, deterministic mutual-information / interpolation step on the sums.

Nothing here needs a GPU except :func:`torch_allreduce`, which wraps the device pointer the trainer hands to
its all-reduce callback (gmg_allreduce_fn, include/gmg_icm.h) in a tensor and calls ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def round_robin(n_items, rank, world):
    """Indices of the items (training strings) rank ``rank`` owns: rank, rank + world, ..."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return np.arange(rank, n_items, world, dtype=np.int64)


def balanced_ranges(offsets, world):
    """Split sequences [0, n) into ``world`` contiguous ranges of (nearly) equal total bases.

    ``offsets`` is the int64 prefix array (n + 1 entries) of the concatenated batch.  Returns an int64 array
    ``cut`` of world + 1 sequence indices, ``cut[r] .. cut[r+1]`` being rank r's range; every sequence lands
    in exactly one range and ranges keep input order, so concatenating per-rank outputs restores it."""
    off = np.asarray(offsets, np.int64)
    n = len(off) - 1
    if world < 1:
        raise ValueError("world must be >= 1")
    total = int(off[-1]) - int(off[0])
    cut = np.zeros(world + 1, np.int64)
    for r in range(1, world):
        target = int(off[0]) + (total * r) // world
        # first sequence whose start is >= target (a sequence is never split)
        cut[r] = int(np.searchsorted(off[:-1], target, side="left"))
    cut[world] = n
    return np.maximum.accumulate(cut)


def take_sequences(ascii_arr, offsets, idx):
    """Gather the sequences ``idx`` of a concatenated batch -> (uint8 ascii, int64 offsets)."""
    off = np.asarray(offsets, np.int64)
    idx = np.asarray(idx, np.int64)
    lens = off[idx + 1] - off[idx]
    new_off = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens, out=new_off[1:])
    if len(idx) and (lens == lens[0]).all() and lens[0] > 0:
        # equal-length strings (the usual training set): one strided gather
        L = int(lens[0])
        src = (off[idx][:, None] + np.arange(L, dtype=np.int64)[None, :]).reshape(-1)
        return np.asarray(ascii_arr, np.uint8)[src], new_off
    out = np.empty(int(new_off[-1]), np.uint8)
    a = np.asarray(ascii_arr, np.uint8)
    for k, i in enumerate(idx):
        out[new_off[k]:new_off[k + 1]] = a[off[i]:off[i + 1]]
    return out, new_off


def slice_range(ascii_arr, offsets, lo, hi):
    """Sequences [lo, hi) of a concatenated batch, as views (no copy) -> (uint8 ascii, int64 offsets)."""
    off = np.asarray(offsets, np.int64)
    a, b = int(off[lo]), int(off[hi])
    return np.asarray(ascii_arr, np.uint8)[a:b], off[lo:hi + 1] - a


def group_by_model(model_of_read):
    """Job list of a many-model run: {model id: indices of its reads, in input order} (ICM-major order)."""
    m = np.asarray(model_of_read)
    order = np.argsort(m, kind="stable")
    ids, starts = np.unique(m[order], return_index=True)
    ends = list(starts[1:]) + [len(m)]
    return {int(i): order[s:e] for i, s, e in zip(ids, starts, ends)}


class _DeviceInt32:
    """__cuda_array_interface__ view of ``count`` int32 at a raw device address."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2}


def torch_allreduce(device_index=None, group=None):
    """An ``allreduce(dptr, count, stream)`` for ICMTraining.Train_Model: sums the level's count slab in place
    across the ranks of ``group`` with torch.distributed, ordered on the trainer's stream."""
    import torch
    import torch.distributed as dist

    def allreduce(dptr, count, stream):
        dev = torch.cuda.current_device() if device_index is None else device_index
        t = torch.as_tensor(_DeviceInt32(dptr, count), device=torch.device("cuda", dev))
        s = torch.cuda.ExternalStream(int(stream), device=dev) if stream else torch.cuda.current_stream(dev)
        with torch.cuda.stream(s):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return allreduce
