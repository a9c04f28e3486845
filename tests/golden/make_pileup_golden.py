#!/usr/bin/env python
"""tests/golden/make_pileup_golden.py -- golden outputs of the REFERENCE's own make_pileup (pileup.pyx:51-113).

    python tests/golden/make_pileup_golden.py          # needs oracle/_ref_pipeline (python oracle/build_ref_pipeline.py)

For a handful of seeded loci of tests/loci.py (one per kind, plus odd settings) the unmodified reference -- compiled against the
stub pysam, its cdef make_pileup reached through oracle/ref_pileup_shim.pyx -- builds its pileup; every read dict is written
to tests/golden/pileup_dicts.json.gz (the `read` object left out, Variant objects as (chrom, pos, ref, alt), qualities as
lists).  tests/test_pileup_ingest.py::test_make_pileup_reproduces_golden replays them through the native ingest without needing
the reference build.  The loci themselves are regenerated from their specs (tests/loci.py is deterministic)."""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import loci as L  # noqa: E402
import refpipe  # noqa: E402

CASES = [
    dict(spec=dict(seed=9001, kind="del", ev_len=3, n_reads=60), excl=True, thresh=20, down=1000),
    dict(spec=dict(seed=9002, kind="ins", ev_len=7, n_reads=60, n_rate=0.01, low_qual_rate=0.08), excl=True, thresh=30, down=1000),
    dict(spec=dict(seed=9003, kind="complex", ev_len=6, ins_len=3, n_reads=60), excl=False, thresh=20, down=1000),
    dict(spec=dict(seed=9004, kind="spliced", ev_len=2, n_reads=80), excl=True, thresh=20, down=1000),
    dict(spec=dict(seed=9005, kind="spliced_hidden", ev_len=5, n_reads=80, low_qual_rate=0.05), excl=True, thresh=25, down=1000),
    dict(spec=dict(seed=9006, kind="hidden_ins", ev_len=11, n_reads=60, clip_frac=0.6, clip_flank=40), excl=True, thresh=20, down=1000),
    dict(spec=dict(seed=9007, kind="del", ev_len=2, n_reads=120, repeat_unit="AC"), excl=True, thresh=20, down=40),        # down-sampled
    dict(spec=dict(seed=9008, kind="long_ins", ev_len=24, n_reads=50, read_len=250, window=167, genome_len=6000, pos=3000), excl=True, thresh=20, down=1000),
]


def prepare(case):
    """the locus of a case exactly as both sides see it (flags planted, coordinate order)"""
    locus = L.make_locus(**case["spec"])
    rng = random.Random(case["spec"]["seed"])
    for r in locus["reads"]:
        r["is_duplicate"] = rng.random() < 0.05
        r["is_secondary"] = rng.random() < 0.03
    locus["reads"].sort(key=lambda r: r["reference_start"])
    return locus


def plain(d):
    out = {}
    for k, v in d.items():
        if k == "read":
            continue
        if k in ("I", "D"):
            out[k] = [[list(x) if hasattr(x, "typecode") else x for x in t[:-1]] + [[t[-1].chrom, t[-1].pos, t[-1].ref, t[-1].alt]] for t in v]
        elif hasattr(v, "typecode"):
            out[k] = list(v)
        elif isinstance(v, tuple):
            out[k] = list(v)
        else:
            out[k] = v
    return out


def main():
    indelpost = refpipe.load()[0]
    import refshim
    from indelpost.local_reference import UnsplicedLocalReference

    doc = []
    for case in CASES:
        locus = prepare(case)
        fa, bam = refpipe.open_locus(locus)
        v = indelpost.Variant(locus["chrom"], locus["pos"], locus["ref"], locus["alt"], fa)
        w = locus["kwargs"]["window"]
        u = UnsplicedLocalReference(v.chrom, v.pos, fa.get_reference_length(v.chrom), w, fa)
        random.seed(99)
        pileup, sf = refshim.ref_make_pileup(v, bam, u, case["excl"], w, case["down"], case["thresh"])
        doc.append(dict(case=case, window=w, rpos=max(x.pos for x in v.generate_equivalents()), sample_factor=sf, pileup=[plain(d) for d in pileup]))
    path = os.path.join(HERE, "pileup_dicts.json.gz")
    with gzip.open(path, "wt") as fh:
        json.dump(doc, fh)
    print("wrote", path, os.path.getsize(path), "bytes;", sum(len(d["pileup"]) for d in doc), "read dicts")


if __name__ == "__main__":
    main()
