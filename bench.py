#!/usr/bin/env python
"""bench.py -- ICM-scored Gbp/s of the B200 hot path, next to the CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Workload (default ``contig5m`` = BASELINE.json configs[1]): one synthetic 5 Mbp bacterial contig per GPU
(seed 20261017 + rank), gene model tests/golden/NC_000915.icm, the
glimmer3 whole-genome scoring half: K1 six-frame ICM walks of every base + per-ORF start enumeration
(Score_Orfs, glimmer3.cc:1275).  A "step" is one pass of that path over the contig.

  value  whole-job Gbp/s with the packed contig and its ORF table already resident in HBM
         (gmg_score_orfs_g3: K1 + K3 count/scan/write), CUDA events, L2 flushed before every step.
  e2e    the same metric through the C-ABI call sequence a host makes with HOST buffers: pinned ASCII ->
         gmg_seqset_create (H2D + Filter/2-bit pack) -> gmg_find_orfs -> gmg_score_orfs_g3 ->
         gmg_get_orfs + gmg_get_starts (D2H), copies inside the timed region.
  roofline   K1 (k1_planes), the dominant kernel: algorithmic HBM bytes per launch
             (0.25 B/base packed read + 6 planes x 4 B/base written = 24.25 B/base) divided by
             its CUDA-event duration measured live in the timed region (gmg_ctx_profile).
  cpu_baseline   the unmodified reference glimmer3 binary (oracle/_ref, built from /root/reference by
             oracle/Makefile) on the same contig on one host core (it is single-threaded).

``--impl reference`` times the reference's own CPU implementation with every host core: one glimmer3
process per core, each on its own bounded slice of the workload (the reference's only parallelism is
process-per-sequence-set, scripts/train_all.py:58).  Multi-GPU: contigs are independent, so ranks shard
them with no collective ("weak" scaling: one contig per GPU).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "ICM-scored Gbp/s"
UNIT = "Gbp/s"
CONTIG_LEN = 5_000_000
K1_BYTES_PER_BASE = 0.25 + 6 * 4  # packed read + six float planes written


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="contig5m", choices=["contig5m"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, units):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, scaled to this launch's units."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fp:
            t = json.load(fp)[kernel]
        return t["dram_bytes_per_unit"] * units
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ref_bin(name):
    p = os.path.join(ROOT, "oracle", "_ref", "bin", name)
    return p if os.path.exists(p) else None


def run_glimmer3_procs(fastas, workdir):
    """Run one reference glimmer3 per FASTA concurrently; returns wall seconds."""
    import workloads as W
    exe = ref_bin("glimmer3")
    t0 = time.perf_counter()
    procs = [subprocess.Popen([exe, "-u", "-12", "-m", W.gene_model_path(), fa, os.path.join(workdir, f"out{i}")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=workdir)
             for i, fa in enumerate(fastas)]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference glimmer3 failed")
    return time.perf_counter() - t0


def port_scoring_seconds(contig_bytes):
    """The oracle port (oracle/icm_oracle.c): Find_Orfs + Score_Orfs start enumeration only, 1 thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import workloads as W
    og = O.lib().orc_icm_read(W.gene_model_path().encode())
    s = bytes(contig_bytes)
    gc = (s.count(b"c") + s.count(b"g")) / len(s)
    oi = O.build_indep(gc)
    op = O.params(False)
    t0 = time.perf_counter()
    orfs = O.find_orfs(s, op)
    O.g3_score_orfs(og, oi, s, op, orfs)
    return time.perf_counter() - t0


def cpu_baseline(contig_arr):
    """cpu_baseline leg of the b200 arm: the whole contig, one reference process (single-threaded binary)."""
    import workloads as W
    tmp = tempfile.mkdtemp(prefix="gmg_bench_")
    try:
        if ref_bin("glimmer3"):
            fa = os.path.join(tmp, "contig.fa")
            W.write_fasta(fa, contig_arr, prefix="contig")
            sec = run_glimmer3_procs([fa], tmp)
            out = {"value": len(contig_arr) / sec / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"the full {len(contig_arr)} bp contig through the unmodified glimmer3 binary "
                             f"(-u -12 -m; FASTA read + Find_Orfs + Score_Orfs + event DP + .predict), {sec:.2f} s"}
            n = min(len(contig_arr), 1_000_000)
            psec = port_scoring_seconds(contig_arr[:n].tobytes())
            out["port_scoring_only"] = {"value": n / psec / 1e9, "unit": UNIT, "cores": 1,
                                        "sample": f"oracle port, Find_Orfs + Score_Orfs start lists only, first {n} bp"}
            return out
        n = min(len(contig_arr), 2_000_000)
        sec = port_scoring_seconds(contig_arr[:n].tobytes())
        return {"value": n / sec / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"oracle port (Find_Orfs + Score_Orfs start lists), first {n} bp of the contig, {sec:.2f} s"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads as W
    cores = host_cores()
    slice_len = 500_000
    tmp = tempfile.mkdtemp(prefix="gmg_ref_")
    try:
        kind = "reference" if ref_bin("glimmer3") else "port"
        n_contigs = max(1, args.gpus)
        contigs = [W.contig(W.CONTIG_SEED + r, CONTIG_LEN) for r in range(n_contigs)]
        slices = []
        for i in range(cores):  # distinct slices, round-robin over the job's contigs
            c = contigs[i % n_contigs]
            a = ((i // n_contigs) * slice_len) % (len(c) - slice_len)
            slices.append(c[a:a + slice_len])
        times = []
        if kind == "reference":
            fastas = []
            for i, s in enumerate(slices):
                fa = os.path.join(tmp, f"slice{i}.fa")
                W.write_fasta(fa, s, prefix="slice")
                fastas.append(fa)
            for k in range(args.warmup + args.steps):
                sec = run_glimmer3_procs(fastas, tmp)
                if k >= args.warmup:
                    times.append(sec)
        else:
            cores = 1
            for k in range(args.warmup + args.steps):
                sec = port_scoring_seconds(slices[0].tobytes())
                if k >= args.warmup:
                    times.append(sec)
        bases = slice_len * (len(slices) if kind == "reference" else 1)
        total = sum(times)
        value = bases * len(times) / total / 1e9
        sample = (f"each step: {cores} concurrent glimmer3 processes (one per host core), each on its own "
                  f"{slice_len} bp slice of the workload contig(s)" if kind == "reference" else
                  f"each step: oracle port on one {slice_len} bp slice, 1 thread")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": "contig5m: synthetic 5 Mbp bacterial contig per GPU, glimmer3 whole-genome "
                                       "scoring (BASELINE.json configs[1])", "model": "tests/golden/NC_000915.icm"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import glimmer_mg_b200 as g
    import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    K, Wu = args.steps, max(args.warmup, 3)
    with torch.cuda.stream(stream):
        ctx = g.Context(local, stream.cuda_stream)
        contig = W.contig(W.CONTIG_SEED + rank, CONTIG_LEN)
        n = len(contig)
        h_ascii = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_ascii.numpy()[:] = contig
        off = np.array([0, n], np.int64)
        gene = g.ICM.Read(ctx, W.gene_model_path())
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

        # ---- resident phase: packed contig + ORF table live in HBM ----
        ss = g.SeqSet(ctx, ascii=h_ascii.numpy(), offsets=off)
        gc = ss.gc_fraction()
        p = g.Params(False)
        p.set_ignore_score_len(gc)
        indep = g.ICM.Build_Indep_WO_Stops(ctx, gc, p.stop_codons)
        n_orfs = ss.find_orfs(p)
        for _ in range(Wu):
            ss.score_orfs_g3(gene, indep, p)
        ctx.sync()
        ctx.profile(True)
        for k in ("k1", "k3", "orf", "pack"):
            ctx.profile_read(k)
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        clocks = ClockSampler(local) if rank == 0 else None
        barrier()
        launches0 = ctx.launches
        t_wall0 = time.time()
        for k in range(K):
            flush.zero_()
            e0[k].record(stream)
            n_starts = ss.score_orfs_g3(gene, indep, p)
            e1[k].record(stream)
        barrier()
        t_wall1 = time.time()
        launches = ctx.launches - launches0
        ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
        k1_ms, k1_n = ctx.profile_read("k1")
        k3_ms, k3_n = ctx.profile_read("k3")
        clk = clocks.stop(t_wall0, t_wall1) if clocks else None

        # ---- end to end: host ASCII in, ORFs + start lists out ----
        e2e_ms = 0.0
        d2h = 0
        for k in range(Wu + K):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            s2 = g.SeqSet(ctx, ascii=h_ascii.numpy(), offsets=off)
            s2.find_orfs(p)
            s2.score_orfs_g3(gene, indep, p)
            orfs, ooff = s2.get_orfs(pinned=True)
            starts, soff = s2.get_starts(pinned=True)
            b.record(stream)
            torch.cuda.synchronize()
            if k >= Wu:
                e2e_ms += a.elapsed_time(b)
                d2h = orfs.nbytes + ooff.nbytes + starts.nbytes + soff.nbytes
            s2.close()
        ctx.profile(False)

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = t.tolist()
    if rank == 0:
        total_bases = n * world
        value = total_bases * K / (ms_max / 1e3) / 1e9
        e2e = total_bases * K / (e2e_max / 1e3) / 1e9
        peak, peak_src = measured_peak()
        k1_bytes = n * K1_BYTES_PER_BASE
        achieved = k1_bytes / (k1_ms / max(k1_n, 1) / 1e3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wu,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "contig5m: synthetic 5 Mbp bacterial contig per GPU (seed 20261017+rank), "
                                       "glimmer3 whole-genome scoring half: K1 six-frame walks + per-ORF start "
                                       "enumeration (BASELINE.json configs[1])",
                           "model": "tests/golden/NC_000915.icm (12/7/3)", "bases_per_gpu": n, "orfs": int(n_orfs),
                           "starts": int(n_starts), "l2": "256 MB flush write before every timed step",
                           "sharding": "one contig per GPU, no collective"},
                "roofline": {"bound": "hbm", "kernel": "k1_planes", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu_traffic("k1_planes", n),
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                             "kernel_ms": k1_ms / max(k1_n, 1), "kernel_share_of_step": k1_ms / ms,
                             "k3_ms_per_step": k3_ms / K},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(n + off.nbytes),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_max / K},
                "gpu_launches": int(launches), "clocks": clk}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(contig)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
