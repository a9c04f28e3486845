"""Record the REAL indelPost call stream: run the unmodified reference pipeline (VariantAlignment + count_alleles + phase,
oracle/_ref_pipeline built by oracle/build_ref_pipeline.py from /root/reference + the stub pysam) on the synthetic loci of
tests/loci.py::parity_specs() and log every Smith-Waterman call it issues with the result the reference's own sswpy/ssw.c
returned.  Output: tests/golden/pipeline_calls.json.gz -- per locus the spec, the pipeline's outputs and the number of calls;
the calls themselves de-duplicated (the pipeline repeats many alignments, pileup.pyx:849 vs 647).

Run in the build container only (needs /root/reference for the build step):
    python oracle/build_ref_pipeline.py && python tests/golden/make_pipeline_golden.py
"""
import gzip
import json
import os
import sys
from collections import Counter

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import loci  # noqa: E402
import refpipe  # noqa: E402


def main():
    seq_table, seq_index = [], {}

    def sid(s):
        s = s.decode() if isinstance(s, bytes) else s
        if s not in seq_index:
            seq_index[s] = len(seq_table)
            seq_table.append(s)
        return seq_index[s]

    out_loci, out_calls, seen = [], [], set()
    sites = Counter()
    for li, spec in enumerate(loci.parity_specs()):
        lc = loci.make_locus(**spec)
        lc["_read_set"] = {r["query_sequence"] for r in lc["reads"]}
        calls = []
        summary = refpipe.run_locus(lc, calls=calls)
        for c in calls:
            site = refpipe.classify_call(c, lc)
            sites[site] += 1
            ref, read, ms, mm, go, ge, s0, e0, out = c
            key = (sid(ref), sid(read), ms, mm, go, ge, s0, e0)
            if key in seen:
                continue
            seen.add(key)
            out_calls.append(dict(locus=li, site=site, ref=key[0], read=key[1], match=ms, mismatch=mm, go=go, ge=ge, start_idx=s0, end_idx=e0, out=list(out)))
        out_loci.append(dict(spec=spec, summary=json.loads(json.dumps(summary)), n_calls=len(calls)))
        print(f"locus {li} {spec['kind']} ev={spec['ev_len']}: counts={summary['counts']} phased={summary['phased']} SW calls={len(calls)}")
    doc = dict(note="every distinct SW call issued by the unmodified reference pipeline on tests/loci.py::parity_specs(), with the reference's own results",
               loci=out_loci, seqs=seq_table, calls=out_calls, sites=dict(sites))
    path = os.path.join(HERE, "pipeline_calls.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(doc, separators=(",", ":")).encode())
    print(f"{path}: {len(out_calls)} distinct calls of {sum(l['n_calls'] for l in out_loci)}, {len(seq_table)} sequences, {os.path.getsize(path)} bytes; sites {dict(sites)}")


if __name__ == "__main__":
    main()
