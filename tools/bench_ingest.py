"""Pileup ingestion rate, host only: reads per second from a BAM file to what the next stage needs,

   reference   : the reference's own make_pileup (pileup.pyx:51-113, through oracle/ref_pileup_shim) on the in-memory stub
                 pysam -- no BAM decoding at all in this arm (pysam / htslib are absent), so it is a LOWER bound of its cost;
   dicts       : indelpost_b200.pileup.make_pileup -- BGZF inflate + BAM decode + index query + dictize_read's integer core in C,
                 the same list of read dicts assembled in Python;
   columnar    : indelpost_b200.pileup.make_pileup_batch + read_table() -- the same ingest without per-read Python objects,
                 ending in the SWB_SEQ_PACKED4 read table swb_align_batch takes;
   fetch       : the region fetch alone (inflate + decode into columns).

    python tools/bench_ingest.py            # one JSON line; bench.py adds it as extra.ingest
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {
    "cfg3_locus_500x150": dict(n_reads=500, read_len=150, window=50, genome_len=4000, pos=2000, downsample=1000),
    "cfg4_locus_20000x250": dict(n_reads=20000, read_len=250, window=167, genome_len=6000, pos=3000, downsample=20000),
}


def _rate(fn, n_reads, budget=1.0):
    fn()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < budget:
        fn(); k += 1
    dt = (time.perf_counter() - t0) / k
    return {"ms": dt * 1e3, "reads_per_s": n_reads / dt}


def measure(budget=1.0):
    import loci as L
    import refpipe
    from indelpost_b200 import bamio, pileup

    have_ref = refpipe.available() and any(f.startswith("refshim") for f in os.listdir(refpipe.REF_PIPELINE))
    out = {"what": "reads/s of pileup ingestion on ONE host core; the reference arm reads in-memory stub records (no BAM decode), ours read BAM + BAI files"}
    tmp = tempfile.mkdtemp(prefix="swb_ingest_")
    for name, sh in SHAPES.items():
        locus = L.make_locus(77, kind="del", ev_len=3, n_reads=sh["n_reads"], read_len=sh["read_len"], window=sh["window"], genome_len=sh["genome_len"], pos=sh["pos"])
        locus["reads"].sort(key=lambda r: r["reference_start"])
        bam_p, fa_p = os.path.join(tmp, name + ".bam"), os.path.join(tmp, name + ".fa")
        bamio.write_fasta(fa_p, {"chr1": locus["genome"]})
        bamio.write_bam(bam_p, [("chr1", len(locus["genome"]))], locus["reads"])
        bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)
        w, ds = sh["window"], sh["downsample"]
        res = {"reads": sh["n_reads"], "bam_bytes": os.path.getsize(bam_p)}
        if have_ref:
            indelpost = refpipe.load()[0]
            import refshim
            from indelpost.local_reference import UnsplicedLocalReference as RefULR

            fa_s, bam_s = refpipe.open_locus(locus)
            v = indelpost.Variant("chr1", locus["pos"], locus["ref"], locus["alt"], fa_s)
            u = RefULR("chr1", v.pos, len(locus["genome"]), w, fa_s)
            res["reference"] = _rate(lambda: refshim.ref_make_pileup(v, bam_s, u, True, w, ds, 20), sh["n_reads"], budget)
            equivalents = v.generate_equivalents
        else:
            class _V:
                pos = locus["pos"]
            equivalents = lambda: [_V]  # noqa: E731

        class Target:
            chrom, pos, reference = "chr1", locus["pos"], fa
            generate_equivalents = staticmethod(equivalents)

        u2 = pileup.UnsplicedLocalReference("chr1", locus["pos"], len(locus["genome"]), w, fa)
        res["dicts"] = _rate(lambda: pileup.make_pileup(Target, bam, u2, True, w, ds, 20), sh["n_reads"], budget)
        res["columnar"] = _rate(lambda: pileup.make_pileup_batch(Target, bam, u2, True, w, ds, 20).read_table(), sh["n_reads"], budget)
        res["fetch"] = _rate(lambda: bam.fetch_columns("chr1", locus["pos"] - 1 - w, locus["pos"] + w), sh["n_reads"], budget)
        if "reference" in res:
            res["columnar_vs_reference"] = res["columnar"]["reads_per_s"] / res["reference"]["reads_per_s"]
            res["dicts_vs_reference"] = res["dicts"]["reads_per_s"] / res["reference"]["reads_per_s"]
        out[name] = res
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
