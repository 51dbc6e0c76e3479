#!/usr/bin/env python3
"""glimmer-mg.py -- front end of the Glimmer-MG pipeline over the B200 binaries.

Keeps the command-line surface of the reference's pipeline driver (/root/reference/scripts/glimmer-mg.py:142-193:
pipeline, Glimmer, Phymm and PhyScimm option groups, same names, defaults and meanings) and builds the same
`glimmer-mg` / `build-icm` command lines (:468-478, :613-626, :51-59), but runs them through this package's hosts:
``glimmer_mg_b200/host/bin/glimmer-mg-gmg`` (the reference driver bound to the GPU scoring path) and
``glimmer_mg_b200/host/bin/build-icm``.  The stages that belong to third-party tools (Phymm classification,
PhyScimm clustering, ELPH / train_features re-training) are invoked exactly where the reference invokes them and are
looked up under ``--tools_dir``; when one is absent the front end stops with the name of the missing program
instead of guessing, and the text-processing stages of the reference script that are not re-implemented here (parsing
raw Phymm output, splitting and merging per-cluster files) stop the same way, naming the reference lines.  ``--dry_run`` prints the commands of every stage without running anything.

This is glue: the hot path lives in libgmgicm.so.  Written for Python 3 (the reference script is Python 2).
"""
import argparse
import glob
import os
import shlex
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "bin")


def build_parser():
    ap = argparse.ArgumentParser(usage="%(prog)s [options] <fasta file>", description="Run the Glimmer-MG pipeline.")
    ap.add_argument("fasta")
    # ---- pipeline (reference :146-165) ----
    ap.add_argument("--iter", dest="iterate", type=int, default=1,
                    help="re-train on the initial predictions and predict again this many times [1]")
    ap.add_argument("--long_orfs", action="store_true", help="make the first ICM with long-orfs")
    ap.add_argument("-o", dest="out", help="prefix of the output files [fasta file name]")
    ap.add_argument("-p", dest="proc", type=int, default=1, help="processes for the classification / clustering tools [1]")
    ap.add_argument("--single_cluster", action="store_true", help="treat all sequences as one cluster (no PhyScimm)")
    ap.add_argument("-t", dest="top_hits", type=int, default=3, help="top Phymm classifications used for training [3]")
    ap.add_argument("--filter", dest="filter_t", type=float, default=1.0, help=argparse.SUPPRESS)
    ap.add_argument("--glim_bin", default=os.path.join(BIN, "glimmer-mg-gmg"), help=argparse.SUPPRESS)
    ap.add_argument("--ignore", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--all_features", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--time", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--skip_first", action="store_true", help=argparse.SUPPRESS)
    # ---- Glimmer (reference :170-176) ----
    g = ap.add_argument_group("Glimmer")
    g.add_argument("-i", "--indel", action="store_true", help="indel mode: predictions may shift the coding frame")
    g.add_argument("-q", dest="quality_file", help="FASTA file of Phred quality values matching the sequences")
    g.add_argument("-r", "--circular", action="store_true", help="circular genome (not supported by the GPU path)")
    g.add_argument("-s", "--sub", action="store_true", help="substitution mode: pass through a mutated stop codon")
    g.add_argument("-u", "--fudge", type=float, default=1.0, help="added to the log-likelihood ratio of every ORF [1.0]")
    # ---- Phymm (reference :182-184) ----
    ph = ap.add_argument_group("Phymm")
    ph.add_argument("--raw", dest="raw_done", action="store_true", help="raw Phymm output exists already")
    ph.add_argument("--class", dest="class_done", action="store_true", help="<prefix>.class.txt exists already")
    # ---- PhyScimm (reference :190-192) ----
    sc = ap.add_argument_group("PhyScimm")
    sc.add_argument("--clust", dest="clust_done", action="store_true", help="cluster FASTA files exist already")
    sc.add_argument("--taxlevel", default="family", help="taxonomic level of the clustering [family]")
    sc.add_argument("--minbp_pct", type=float, default=0.01, help="minimum share of bases for a class to become a cluster [0.01]")
    # ---- this front end ----
    x = ap.add_argument_group("front end")
    x.add_argument("--tools_dir", default=os.environ.get("GLIMMER_MG_TOOLS", ""),
                   help="directory holding the third-party stages (phymm_par.py, physcimm.py, train_features.py, long-orfs, extract)")
    x.add_argument("--dry_run", action="store_true", help="print the commands of every stage, run nothing")
    return ap


class Runner:
    def __init__(self, opts):
        self.opts = opts
        self.log = []

    def tool(self, name):
        for d in (self.opts.tools_dir, BIN):
            if d and os.path.exists(os.path.join(d, name)):
                return os.path.join(d, name)
        if self.opts.dry_run:
            return name
        sys.exit(f"glimmer-mg.py: stage needs `{name}`, which is not part of this package; point --tools_dir "
                 f"(or $GLIMMER_MG_TOOLS) at a directory that holds it")

    def pipeline_stage(self, what, where):
        """A text-processing stage of the reference's pipeline script that this package does not re-implement (SURVEY.md
        section 2 row 14: option surface kept, pipeline logic out of scope)."""
        line = f"# reference pipeline stage: {what} ({where})"
        self.log.append(line)
        if self.opts.dry_run:
            print(line)
            return
        sys.exit(f"glimmer-mg.py: {what} is a stage of the reference's pipeline script ({where}) that this package does not "
                 f"provide; run that stage with the reference's scripts and pass its result (--class / --clust)")

    def run(self, argv, stdin=None, stdout=None):
        line = " ".join(shlex.quote(a) for a in argv)
        if stdin:
            line += " < " + shlex.quote(stdin)
        if stdout:
            line += " > " + shlex.quote(stdout)
        self.log.append(line)
        if self.opts.dry_run:
            print(line)
            return
        fin = open(stdin, "rb") if stdin else None
        fout = open(stdout, "wb") if stdout else None
        try:
            rc = subprocess.call(argv, stdin=fin, stdout=fout)
        finally:
            for f in (fin, fout):
                if f:
                    f.close()
        if rc != 0:
            sys.exit(f"glimmer-mg.py: `{line}` failed with exit status {rc}")


def glimmer_command(opts):
    """`glimmer-mg -u <fudge> [-i] [-r] [-s]` (reference :468-478)."""
    cmd = [opts.glim_bin, "-u", "%f" % opts.fudge]
    for flag, on in (("-i", opts.indel), ("-r", opts.circular), ("-s", opts.sub)):
        if on:
            cmd.append(flag)
    return cmd


def repredict(R, g3, fasta, prefix, class_file, qual):
    """Re-train on the previous iteration's predictions and predict again (reference :613-660)."""
    o = R.opts
    for it in range(2, o.iterate + 2):
        prev = f"{prefix}.run{it - 1}"
        nxt = f"{prefix}.run{it}" if it < o.iterate else prefix
        if not o.dry_run:
            keep_good_predictions(prev + ".predict", prev + ".fpredict", o.filter_t)
        R.run([R.tool("train_features.py"), "-f", *(["--indel"] if o.indel else []), "--seq", fasta, "--predict", prev + ".fpredict"])
        if not o.all_features and not o.dry_run:
            keep_start_distribution(prev + ".features.txt")
        R.run([*g3, "-b", prev + ".motif", "-m", prev + ".gicm", "-f", prev + ".features.txt", "-c", class_file, *qual, fasta, nxt])


def keep_good_predictions(src, dst, threshold):
    """Predictions that score above the threshold feed the re-training (reference filter_predictions)."""
    with open(src) as fin, open(dst, "w") as fout:
        for line in fin:
            if line.startswith(">") or float(line.split()[4]) >= threshold:
                fout.write(line)


def keep_start_distribution(path):
    """Only the start-codon distribution of a features file unless --all_features (reference :648-660)."""
    keep, on = [], False
    for line in open(path):
        if line.startswith("DIST"):
            on = line.startswith("DIST START")
        if on:
            keep.append(line)
    with open(path, "w") as f:
        f.writelines(keep)


def main(argv=None):
    opts = build_parser().parse_args(argv)
    t_all = time.time()
    fasta = opts.fasta
    prefix = opts.out or os.path.splitext(os.path.basename(fasta))[0]
    R = Runner(opts)
    if opts.circular:
        sys.exit("glimmer-mg.py: -r / --circular: the GPU scoring path handles linear sequences only")
    class_file = prefix + ".class.txt"
    if not opts.class_done:
        if not opts.raw_done:  # Phymm classification (reference :41-46)
            R.run([R.tool("phymm_par.py"), "-b", "-p", str(opts.proc), fasta])
        R.pipeline_stage(f"parse the raw Phymm scores into {class_file} (top {opts.top_hits} hits)", "scripts/glimmer-mg.py:48-57, parse_phymm")
    elif opts.iterate != 0 and not opts.single_cluster and not opts.raw_done:
        sys.exit("glimmer-mg.py: cannot use --class for multiple iterations with clustering: the Phymm scores are needed")
    icm = []
    if opts.long_orfs:  # first ICM from long ORFs (reference :51-59)
        R.run([R.tool("long-orfs"), "-n", "-t", "1.15", fasta, prefix + ".longorfs"])
        R.run([R.tool("extract"), "-t", fasta, prefix + ".longorfs"], stdout=prefix + ".train")
        model = prefix + (".icm" if opts.iterate == 0 else ".run1.icm")
        R.run([os.path.join(BIN, "build-icm"), "-r", model], stdin=prefix + ".train")
        icm = ["-m", model]
    g3 = glimmer_command(opts)
    qual = ["-q", opts.quality_file] if opts.quality_file else []
    t0 = time.time()
    if opts.iterate == 0:
        R.run([*g3, *icm, "-c", class_file, *qual, fasta, prefix])
    else:
        if not opts.skip_first:
            R.run([*g3, *icm, "-c", class_file, *qual, fasta, prefix + ".run1"])
        if opts.single_cluster:
            repredict(R, g3, fasta, prefix, class_file, qual)
        else:
            if not opts.clust_done:  # PhyScimm clustering (reference :104-107)
                R.run([R.tool("physcimm.py"), "-s", fasta, "-p", str(opts.proc), "-r",
                       "results.01.phymm_%s.txt" % os.path.basename(fasta).replace(".", "_"), "--taxlevel", opts.taxlevel,
                       "--minbp_pct", "%f" % opts.minbp_pct])
            clusters = sorted(glob.glob("cluster*fa"))
            if not clusters and not opts.dry_run:
                sys.exit("glimmer-mg.py: cluster FASTA files not found; drop --clust")
            for cf in clusters:  # one glimmer-mg run per cluster with the cluster's own models (reference :117-131)
                sub = f"{prefix}.{cf[:-3]}"
                R.pipeline_stage(f"split {class_file} and {prefix}.run1.predict for {cf}", "scripts/glimmer-mg.py:269-326, cluster_repredict")
                repredict(R, g3, cf, sub, sub + ".class.txt", qual)
            R.pipeline_stage(f"merge the per-cluster predictions into {prefix}.predict", "scripts/glimmer-mg.py:332-400, combine_predictions")
    if opts.time and not opts.dry_run:
        with open(f"time_{prefix}.txt", "w") as f:
            f.write("%.3fs\n" % (time.time() - t_all))
        with open(f"time_{prefix}_iter0.txt", "w") as f:
            f.write("%.3fs\n" % (time.time() - t0))
    return R.log


if __name__ == "__main__":
    main()
