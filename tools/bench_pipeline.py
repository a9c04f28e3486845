"""Pipeline-level measurement (BASELINE.json configs 1 / 3 / 4): loci per second of the UNMODIFIED reference pipeline
(`VariantAlignment` + `count_alleles` + `phase`, oracle/_ref_pipeline + stub pysam) on synthetic loci,

   reference : on its own sswpy / ssw.c (one host process per worker, loci split evenly),
   wave      : the same control flow with `indelpost.localn.SSW = indelpost_b200.SSW` under the wave scheduler
               (indelpost_b200/wave.py): all loci of a worker advance together, one merged GPU batch per wave,

and the check that both give identical outputs for every locus.  The host logic above the SW calls is the reference's own
single-threaded Python/Cython in both arms (out of scope), so the ratio is Amdahl-bounded by the
SW share of a locus (54-64 %); scaling beyond one host core is by processes (loci are independent).

    python tools/bench_pipeline.py [--config cfg3] [--loci N] [--workers P] [--arms reference,wave] [--devices 0,1,...] [--from-files]

Used by bench.py (`extra.pipeline`) on rank 0 with a bounded sample; standalone it prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # BASELINE.json configs[0]: single 1-bp deletion locus, 200 x 150-bp reads
    "cfg1": dict(kinds=[("del", 1)], n_reads=200, read_len=150, window=50),
    # configs[2]: indel loci x 500 reads (150 bp), mixed indel types (del / ins 1-20 bp, 10 % complex)
    "cfg3": dict(kinds=[("del", 1), ("ins", 2), ("del", 4), ("ins", 6), ("del", 9), ("ins", 12), ("hidden_del", 3), ("hidden_ins", 5), ("del", 17), ("complex", 6)],
                 n_reads=500, read_len=150, window=50),
    # configs[3]: high-depth amplicon, 250-bp reads vs ~1 kb windows (window=167), complex indels; depth bounded for the sample
    "cfg4": dict(kinds=[("complex", 9), ("del", 12), ("ins", 15), ("complex", 14)], n_reads=2000, read_len=250, window=167, genome_len=6000, pos=3000),
}


def make_specs(config: str, n_loci: int, seed0: int = 5000):
    c = CONFIGS[config]
    specs = []
    for k in range(n_loci):
        kind, ev = c["kinds"][k % len(c["kinds"])]
        sp = dict(seed=seed0 + k, kind=kind, ev_len=ev, n_reads=c["n_reads"], read_len=c["read_len"], window=c["window"])
        if kind == "complex":
            sp["ins_len"] = max(2, ev // 2)
        for opt in ("genome_len", "pos"):
            if opt in c:
                sp[opt] = c[opt]
        specs.append(sp)
    return specs


def write_loci_files(lcs, directory, tag="loci"):
    """all loci of a worker into ONE coordinate-sorted BAM + BAI and ONE FASTA + FAI (every locus its own contig), written by
    libswbbam -> (bamio.AlignmentFile, bamio.FastaFile).  BASELINE.json's configs are "written to BAM"; pysam is absent."""
    from indelpost_b200 import bamio

    reads, seqs = [], {}
    for k, lc in enumerate(lcs):
        name = f"locus{k:05d}"
        lc["chrom"] = name
        for r in lc["reads"]:
            r["reference_name"] = name
        seqs[name] = lc["genome"]
        reads.extend(lc["reads"])
    bam_p, fa_p = os.path.join(directory, tag + ".bam"), os.path.join(directory, tag + ".fa")
    bamio.write_fasta(fa_p, seqs)
    bamio.write_bam(bam_p, [(k, len(v)) for k, v in seqs.items()], reads)
    return bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)


def _worker(arm, specs, device, q, from_files=False):
    """one host process: its share of the loci through one arm -> (seconds, summaries, stats)"""
    try:
        import loci
        import refpipe

        lcs = [loci.make_locus(**sp) for sp in specs]          # generation is not part of the timed region
        refpipe.load()
        files = None
        base_bam = refpipe.load()[1].AlignmentFile
        if from_files:
            import tempfile

            tmpdir = tempfile.mkdtemp(prefix="swb_loci_")
            files = write_loci_files(lcs, tmpdir, f"w{os.getpid()}")
            base_bam = refpipe.file_backed_bam(files[0])
        if arm == "reference":
            t0 = time.perf_counter()
            outs = [refpipe.run_locus(lc, files=files) for lc in lcs]
            dt = time.perf_counter() - t0
            q.put((dt, outs, {}))
            return
        from indelpost_b200 import SSW, clear_prefetched, wave
        from indelpost_b200.sswpy import _aligner

        if arm == "percall":
            # ZERO-CHANGE use: the product's SSW class swapped in, nothing else -- no prefetch line, no wave scheduler; every miss
            # is a device round trip (widened by the implicit batching of sswpy.py)
            from indelpost_b200 import sswpy

            refpipe.run_locus(lcs[0], ssw_cls=SSW, files=files)           # warm-up: context, kernels
            clear_prefetched()
            sswpy.auto_stats.update(batches=0, pairs=0, hits=0)
            t0 = time.perf_counter()
            outs = [refpipe.run_locus(lc, ssw_cls=SSW, files=files) for lc in lcs]
            dt = time.perf_counter() - t0
            q.put((dt, outs, {"waves": sswpy.auto_stats["batches"], "pairs": sswpy.auto_stats["pairs"], "requests": sswpy.auto_stats["hits"]}))
            return

        al = _aligner(device)
        tee = wave.tee_alignment_file(base_bam)
        runner = wave.WaveRunner(device=device, aligner=al, max_inflight=256)
        with refpipe.swapped(SSW):
            # warm-up: context, kernels, staging buffers
            runner.map(lambda lc: refpipe.run_locus(lc, swap=False, bam_cls=tee, files=files), lcs[:1])
            runner.stats = {k: 0 for k in runner.stats}
            clear_prefetched()
            t0 = time.perf_counter()
            outs = runner.map(lambda lc: refpipe.run_locus(lc, swap=False, bam_cls=tee, files=files), lcs)
            dt = time.perf_counter() - t0
        q.put((dt, outs, dict(runner.stats)))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((None, traceback.format_exc(), {}))


def run_arm(arm, specs, workers, devices, from_files=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    shards = [specs[k::workers] for k in range(workers)]
    procs = []
    t0 = time.perf_counter()
    for k, sh in enumerate(shards):
        p = ctx.Process(target=_worker, args=(arm, sh, devices[k % len(devices)], q, from_files))
        p.start()
        procs.append(p)
    import queue

    parts = []
    while len(parts) < len(procs):
        try:
            parts.append(q.get(timeout=1.0))
        except queue.Empty:                                   # a worker that died without reporting must not hang the bench
            if all(not p.is_alive() for p in procs) and q.empty():
                raise RuntimeError(f"{arm}: a worker exited without a result (exit codes {[p.exitcode for p in procs]})")
    for p in procs:
        p.join()
    wall = time.perf_counter() - t0
    for dt, outs, st in parts:
        if dt is None:
            raise RuntimeError(f"{arm} worker failed:\n{outs}")
    return parts, wall, shards


# ---- the product's locus-parallel driver (indelpost_b200.locuspool) on ONE BAM + FASTA holding every locus -------------------------
_POOL = {}


def _pool_init(rank, device):
    """in every worker: open the files, load the reference pipeline, install the SSW class of the arm"""
    import types

    import refpipe
    from indelpost_b200 import bamio

    refpipe.load()
    bam, fa = bamio.AlignmentFile(os.environ["SWB_POOL_BAM"]), bamio.FastaFile(os.environ["SWB_POOL_FA"])
    _POOL["files"] = (bam, fa)
    base = refpipe.file_backed_bam(bam)
    arm = os.environ["SWB_POOL_ARM"]
    if arm == "pool_reference":
        _POOL["bam_cls"] = base
        return None
    from indelpost_b200 import SSW, sswpy, wave

    _POOL["bam_cls"] = wave.tee_alignment_file(base)
    if os.environ.get("SWB_POOL_CPU_ORACLE"):            # host-logic check without a GPU: batches computed by the CPU oracle
        import test_wave_cpu as W

        sswpy.align_batch = W._oracle_align_batch([])
        refpipe.load()[2].SSW = W._NoGpuSSW
        return types.SimpleNamespace(aligner=object())
    refpipe.load()[2].SSW = SSW
    return types.SimpleNamespace(aligner=sswpy._aligner(device))


def _pool_item(meta):
    import refpipe

    swap = os.environ["SWB_POOL_ARM"] == "pool_reference"
    return refpipe.run_locus(meta, swap=swap, bam_cls=_POOL["bam_cls"], files=_POOL["files"])


def measure_pool(config="cfg3", n_loci=64, workers=4, devices=(0,), arms=("pool_reference", "pool_wave"), cpu_oracle=False):
    """loci/s through indelpost_b200.locuspool.LocusPool: every locus in ONE coordinate-sorted BAM + BAI and one FASTA + FAI
    (written by libswbbam, not timed), a list of (chrom, pos, ref, alt) work items, `workers` processes over `devices`.
    Timed region: pool.map() over all items in the parent (workers already started and warmed by one item each)."""
    import tempfile

    import loci
    from indelpost_b200 import locuspool

    specs = make_specs(config, n_loci)
    lcs = [loci.make_locus(**sp) for sp in specs]
    tmpdir = tempfile.mkdtemp(prefix="swb_pool_")
    bam, fa = write_loci_files(lcs, tmpdir, "all_loci")
    os.environ["SWB_POOL_BAM"], os.environ["SWB_POOL_FA"] = bam.filename, fa.filename
    bam.close(); fa.close()
    if cpu_oracle:
        os.environ["SWB_POOL_CPU_ORACLE"] = "1"
    metas = [dict(chrom=lc["chrom"], pos=lc["pos"], ref=lc["ref"], alt=lc["alt"], kwargs=lc["kwargs"]) for lc in lcs]
    out = {"config": config, "loci": n_loci, "reads": sum(sp["n_reads"] for sp in specs), "workers": workers, "devices": list(devices),
           "bam_bytes": os.path.getsize(os.environ["SWB_POOL_BAM"]),
           "what": "indelpost_b200.locuspool.LocusPool over one BAM + FASTA holding every locus (native reader in both arms); timed region = pool.map() in the parent"}
    outs = {}
    for arm in arms:
        os.environ["SWB_POOL_ARM"] = arm
        with locuspool.LocusPool(_pool_item, workers=workers, devices=devices, init=_pool_init, mode="plain" if arm == "pool_reference" else "wave") as pool:
            pool.map(metas[:workers])                    # warm-up: contexts, kernels, page cache
            t0 = time.perf_counter()
            res = pool.map(metas)
            dt = time.perf_counter() - t0
            outs[arm] = res
            out[arm] = {"seconds": dt, "loci_per_s": n_loci / dt, "reads_per_s": out["reads"] / dt}
            if pool.stats:
                out[arm].update({"waves": pool.stats.get("waves"), "pairs_aligned": pool.stats.get("pairs")})
    if len(outs) == 2:
        a, b = (outs[k] for k in arms)
        out["identical_outputs"] = a == b
        out["speedup"] = out[arms[1]]["loci_per_s"] / out[arms[0]]["loci_per_s"]
    return out


def measure(config="cfg3", n_loci=32, workers=1, arms=("reference", "wave"), devices=(0,), repeats=1, from_files=False):
    specs = make_specs(config, n_loci)
    n_reads = sum(sp["n_reads"] for sp in specs)
    out = {"config": config, "loci": n_loci, "reads": n_reads, "workers": workers, "devices": list(devices),
           "repeats": repeats, "input": "BAM + BAI / FASTA + FAI files written and read by libswbbam (one BAM per worker, one contig per locus)" if from_files else "in-memory stub pysam",
           "what": "unmodified reference VariantAlignment + count_alleles + phase per locus (stub pysam, synthetic loci); timed region = the pipeline calls, "
                   "max over worker processes; host logic above the SW calls is the reference's own in both arms"}
    by_arm = {}
    for arm in arms:
        parts, wall, shards = run_arm(arm, specs, workers, devices, from_files)
        dt = max(p[0] for p in parts)
        for _ in range(max(0, repeats - 1)):                 # the host is shared: keep the fastest of a few runs
            parts2, wall2, _ = run_arm(arm, specs, workers, devices, from_files)
            dt2 = max(p[0] for p in parts2)
            if dt2 < dt:
                parts, wall, dt = parts2, wall2, dt2
        # results back in locus order: worker k got specs[k::workers]; parts arrive in completion order -> match by content
        summaries = {}
        for p in parts:
            for o in p[1]:
                summaries.setdefault(json.dumps(o, sort_keys=True, default=str), 0)
                summaries[json.dumps(o, sort_keys=True, default=str)] += 1
        by_arm[arm] = summaries
        stats = {}
        for p in parts:
            for k, v in p[2].items():
                stats[k] = stats.get(k, 0) + v
        out[arm] = {"seconds": dt, "loci_per_s": n_loci / dt, "reads_per_s": n_reads / dt, "wall_incl_startup_s": wall}
        if stats:
            out[arm]["waves"] = stats.get("waves")
            out[arm]["pairs_aligned"] = stats.get("pairs")
            out[arm]["requests"] = stats.get("requests")
    if "reference" in by_arm and "wave" in by_arm:
        out["identical_outputs"] = by_arm["reference"] == by_arm["wave"]
        out["speedup"] = out["wave"]["loci_per_s"] / out["reference"]["loci_per_s"]
    if "reference" in by_arm and "percall" in by_arm:
        out["percall_identical_outputs"] = by_arm["reference"] == by_arm["percall"]
        out["percall_vs_reference"] = out["percall"]["loci_per_s"] / out["reference"]["loci_per_s"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--loci", type=int, default=32)
    ap.add_argument("--workers", type=int, default=1)
    ap.add_argument("--arms", default="reference,wave")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--from-files", action="store_true", help="loci are written to BAM / FASTA files and read back through the native reader (indelpost_b200.bamio)")
    ap.add_argument("--pool", action="store_true", help="measure through indelpost_b200.locuspool.LocusPool on one BAM holding every locus")
    ap.add_argument("--cpu-oracle", action="store_true", help="(--pool) merged batches computed by the CPU oracle: checks the host logic without a GPU")
    a = ap.parse_args()
    if a.pool:
        print(json.dumps(measure_pool(a.config, a.loci, a.workers, tuple(int(x) for x in a.devices.split(",")), cpu_oracle=a.cpu_oracle)))
        return
    print(json.dumps(measure(a.config, a.loci, a.workers, tuple(a.arms.split(",")), tuple(int(x) for x in a.devices.split(",")), from_files=a.from_files)))


if __name__ == "__main__":
    main()
