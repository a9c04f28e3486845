"""BAM file -> native columnar ingest -> packed read table -> swb_align_batch on the GPU, against the oracle.

The reads never exist as Python strings on the product side: `make_pileup_batch` fetches the locus' records into columns,
`read_table()` converts the BAM's own 4-bit bases into the SWB_SEQ_PACKED4 table in one C pass, the windows (the +-3 x window
reference slice of UnsplicedLocalReference.fetch_ref_seq, local_reference.pyx:22-30) are packed by swb_pack_table, and every
kept read is aligned under indelPost's six-point penalty grid (varaln.pyx:1127-1143).  The oracle aligns the same pairs built
from the plain strings of the locus generator."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import loci as L  # noqa: E402
import swbtest as T  # noqa: E402

pytestmark = pytest.mark.gpu

GRID = ((3, 1), (3, 0), (4, 1), (4, 0), (5, 1), (5, 0))


def test_gpu_bam_to_batch_equals_oracle(tmp_path):
    from gpuutil import aligner
    from indelpost_b200 import bamio, pileup
    from indelpost_b200.batch import pack_table

    specs = [dict(seed=4100 + k, kind=kind, ev_len=ev, n_reads=160, n_rate=0.005, low_qual_rate=0.02)
             for k, (kind, ev) in enumerate([("del", 3), ("ins", 7), ("complex", 6), ("hidden_del", 4), ("spliced", 2), ("long_ins", 24), ("del", 1), ("ins", 2)])]
    lcs = [L.make_locus(**sp) for sp in specs]
    reads, seqs = [], {}
    for k, lc in enumerate(lcs):
        name = f"locus{k}"
        lc["chrom"] = name
        for r in lc["reads"]:
            r["reference_name"] = name
        seqs[name] = lc["genome"]
        reads.extend(lc["reads"])
    bam_p, fa_p = os.path.join(str(tmp_path), "loci.bam"), os.path.join(str(tmp_path), "loci.fa")
    bamio.write_fasta(fa_p, seqs)
    bamio.write_bam(bam_p, [(k, len(v)) for k, v in seqs.items()], reads)
    bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)

    # product side: columns and packed tables only
    tables, offs, lens, win_ascii, pr, pw, go, ge = [], [], [], [], [], [], [], []
    o_reads, o_pairs = [], []           # oracle side: plain strings
    base_read = base_byte = 0
    for k, lc in enumerate(lcs):
        class Target:
            chrom, pos, reference = lc["chrom"], lc["pos"], fa

            @staticmethod
            def generate_equivalents(p=lc["pos"]):
                return [type("V", (), {"pos": p})]

        u = pileup.UnsplicedLocalReference(lc["chrom"], lc["pos"], len(lc["genome"]), 50, fa)
        pb = pileup.make_pileup_batch(Target, bam, u, True, 50, 1000, 20)
        table, off, ln, keep = pb.read_table()
        tables.append(table); offs.append(off + base_byte); lens.append(ln)
        win_ascii.append(u.fetch_ref_seq(lc["pos"], 50).encode())
        by_name = {r["query_name"]: r["query_sequence"] for r in lc["reads"]}
        for i in keep.tolist():
            o_reads.append(by_name[pb.batch.name(i)])
            for (a, e) in GRID:
                pr.append(base_read + i); pw.append(k); go.append(a); ge.append(e)
                o_pairs.append((len(o_reads) - 1, k))
        base_read += len(off); base_byte += len(table)
    assert len(pr) > 5000
    wblob = np.frombuffer(b"".join(win_ascii), "u1")
    wlen = np.array([len(w) for w in win_ascii], "<i4"); woff = np.concatenate([[0], np.cumsum(wlen[:-1])]).astype("<i8")
    wtab, wtoff = pack_table(wblob, woff, wlen, bits=4, ascii=True)
    mat = T.dna_matrix(3, 2)
    res, arena = aligner().align(np.concatenate(tables).view(np.int8), np.concatenate(offs), np.concatenate(lens), wtab.view(np.int8), wtoff, wlen,
                                 np.array(pr, "<i4"), np.array(pw, "<i4"), np.array(go, "u1"), np.array(ge, "u1"), mat=mat, seq_encoding=2)
    ob = T.batch_from_lists([T.encode_dna(s) for s in o_reads], [T.encode_dna(w.decode()) for w in win_ascii],
                            [p[0] for p in o_pairs], [p[1] for p in o_pairs], np.array(go, "u1"), np.array(ge, "u1"), mat=mat)
    ro, ao = T.oracle_parallel(ob, threads=min(16, os.cpu_count() or 1))
    T.compare(res.view(T.RESULT_DTYPE), arena, ro, ao, what="BAM -> columns -> packed table -> GPU vs oracle on the generator's strings")
    assert int((res["cigar_len"] > 1).sum()) > 500          # gapped alignments are in the mix
