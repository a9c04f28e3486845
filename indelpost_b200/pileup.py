"""Pileup ingestion on the native reader: `make_pileup` of pileup.pyx:51-113 with the same name,
argument meaning and result -- the list of read dicts every later stage of indelPost pattern-matches on -- but with ONE
region fetch into a columnar batch and the per-read integer work (`dictize_read`, pileup.pyx:160-266, and the helpers it calls
in utilities.pyx:187-327, 429-503) done in C over the whole batch (`swb_pileup_columns`, include/swbbam.h).  Python only
assembles the dicts from the columns.  `make_pileup_batch` is the same ingest without dicts: the reads of the locus as a
SWB_SEQ_PACKED4 table ready for `BatchAligner` / `swb_align_batch`, plus the columns.

The reference's `make_pileup` is a `cdef` function (`varaln.pyx:30` cimports it), so it cannot be swapped from Python; the
patch a maintainer adds is in INTEGRATION.md §4.  Parity is pinned against the reference's own `make_pileup`, reached through
a four-line Cython shim (oracle/ref_pileup_shim.pyx), on every locus of tests/loci.py: tests/test_pileup_ingest.py.
"""
from __future__ import annotations

import random
import re
from collections import namedtuple
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import bamio

cigar_ptrn = re.compile(r"[0-9]+[MIDNSHPX=]")

#: what `I` / `D` entries carry as their last element when no Variant class is supplied
IndelCall = namedtuple("IndelCall", "chrom pos ref alt")


def _plain_variant(chrom, pos, ref, alt, reference, skip_validation=True):
    return IndelCall(chrom, pos, ref, alt)


class UnsplicedLocalReference:
    """local_reference.pyx:4-36: the reference window of +-10 x `window` around the target, fetched once"""

    def __init__(self, chrom, pos, ref_len, window, reference):
        self.chrom, self.pos, self.ref_len, self.window = chrom, pos, ref_len, window
        self.local_ref_start = max(0, pos - window * 10)
        self.unspliced_local_reference = reference.fetch(chrom, self.local_ref_start, min(pos + window * 10, ref_len))
        self.left_len = 0

    def fetch_ref_seq(self, target_pos, window):
        self.left_len = target_pos - max(0, target_pos - window * 3)
        return self.get_ref_seq(max(0, target_pos - window * 3), min(target_pos + window * 3, self.ref_len))

    def get_ref_seq(self, start, end):
        i = start - self.local_ref_start
        return self.unspliced_local_reference[i: i + (end - start)]


def bam_chrom(chrom: str, bam) -> str:
    """pileup.pyx:71-78: the BAM may name the contig with or without the `chr` prefix"""
    if chrom in bam.references:
        return chrom
    return chrom.replace("chr", "") if chrom.startswith("chr") else "chr" + chrom


def fetch_reads(chrom, pos, bam, ref_len, window, exclude_duplicates):
    """pileup.pyx:126-157 -> (ReadBatch of the region, index array of the records the reference keeps, in file order)"""
    pos = pos - 1
    batch = bam.fetch_columns(chrom, max(0, pos - window), min(pos + 1 + window, ref_len))
    flag, ncig, start = batch.flag, batch.n_cigar, batch.pos
    keep = (flag & bamio.FSECONDARY) == 0
    keep &= ncig > 0
    if exclude_duplicates:
        keep &= (flag & bamio.FDUP) == 0
        keep &= start != 0          # `and read.reference_start` (pileup.pyx:145): a read at position 0 is dropped
    return batch, keep.nonzero()[0]


def is_within_intron(read, pos, window):
    intron = read["intron_pattern"]
    if intron == (0, 0):
        return False
    return intron[0] < pos - window and pos + window < intron[1]


def _select(target, bam, window, downsamplethresh, exclude_duplicates):
    """fetch + the down-sampling decision of pileup.pyx:80-108 -> (batch, kept record indices, sample_factor, ref_len, rpos)"""
    chrom, pos, reference = target.chrom, target.pos, target.reference
    rpos = max(v.pos for v in target.generate_equivalents())
    ref_len = reference.get_reference_length(chrom)
    _chrom = bam_chrom(chrom, bam)
    batch, keep = fetch_reads(_chrom, pos, bam, ref_len, window, exclude_duplicates)
    # bam.count(_chrom, pos - 1, pos, read_callback=...) (pileup.pyx:82-83): [pos - 1, pos) lies inside the fetched region, so
    # the count is taken from the batch instead of a second pass over the file
    orig_depth = batch.count_overlapping(pos - 1, pos, (bamio.FUNMAP | bamio.FSECONDARY | bamio.FQCFAIL | bamio.FDUP) if exclude_duplicates else 0)
    orig_read_num = len(keep)
    sample_factor = 1.0
    if orig_depth > downsamplethresh:
        random.seed(123)
        n_sample = int(orig_read_num * (downsamplethresh / orig_depth))
        if n_sample >= downsamplethresh / 2 > 0:
            keep = np.array(random.sample(keep.tolist(), n_sample), dtype=np.int64)   # the draw depends on len() only: same reads as sampling the segment list
            sample_factor = orig_read_num / len(keep)
    return batch, keep, sample_factor, ref_len, rpos


def _columns(batch, keep, target, rpos, unspl_loc_ref, basequalthresh):
    """one reference fetch covering every kept read and the local window, then swb_pileup_columns"""
    chrom, reference = target.chrom, target.reference
    lo = unspl_loc_ref.local_ref_start
    hi = lo + len(unspl_loc_ref.unspliced_local_reference)
    if len(keep):
        lo = min(lo, int(batch.pos[keep].min()))
        hi = max(hi, int(batch.end[keep].max()) + 1)
    lo = max(0, lo)
    contig = reference.fetch(chrom, lo, hi)
    return batch.pileup_columns(target.pos, rpos, basequalthresh, contig.encode("ascii"), lo, unspl_loc_ref.local_ref_start,
                                len(unspl_loc_ref.unspliced_local_reference))


def make_pileup(target, bam, unspl_loc_ref, exclude_duplicates, window, downsamplethresh, basequalthresh,
                variant_factory: Optional[Callable] = None) -> Tuple[List[dict], float]:
    """pileup.pyx:51-113.  `bam`: bamio.AlignmentFile; `target`: anything with chrom / pos / reference / generate_equivalents()
    (indelpost.Variant); `unspl_loc_ref`: UnsplicedLocalReference (above); `variant_factory(chrom, pos, ref, alt, reference,
    skip_validation=True)` builds the object stored last in every `I` / `D` entry (pass indelpost.Variant for the
    reference's own type; default: an IndelCall tuple)."""
    make_var = variant_factory or _plain_variant
    chrom, pos, reference = target.chrom, target.pos, target.reference
    batch, keep, sample_factor, ref_len, rpos = _select(target, bam, window, downsamplethresh, exclude_duplicates)
    cols = _columns(batch, keep, target, rpos, unspl_loc_ref, basequalthresh)
    R = {name: cols.reads[name].tolist() for name in cols.reads.dtype.names if name != "pad_"}
    sub = cols.subreads.tolist()
    ind = cols.indels.tolist()
    flag, mapq = batch.flag.tolist(), batch.mapq.tolist()
    pileup = []
    for i in keep.tolist():
        seg = batch.segment(i)
        read_seq, read_qual, cigar_string = seg.query_sequence, seg.query_qualities, seg.cigarstring
        ref_seq = cols.ref_seq(i)
        read_dict = {
            "read": seg,
            "read_seq": read_seq,
            "read_qual": read_qual,
            "ref_seq": ref_seq,
            "is_reverse": bool(flag[i] & bamio.FREVERSE),
            "read_name": seg.query_name,
            "mapq": mapq[i],
            "start_offset": R["start_offset"][i],
            "aln_start": R["aln_start"][i],
            "read_start": R["read_start"][i],
            "end_offset": R["end_offset"][i],
            "aln_end": R["aln_end"][i],
            "read_end": R["read_end"][i],
            "cigar_string": cigar_string,
            "cigar_list": cigar_ptrn.findall(cigar_string),
            "is_reference_seq": bool(R["is_reference_seq"][i]),
            "I": [],
            "D": [],
            "low_qual_base_num": R["low_qual_base_num"][i],
            "is_end_dirty": bool(R["is_end_dirty"][i]),
            "is_dirty": bool(R["is_dirty"][i]),
        }
        spliced_cigar = "N" in cigar_string
        o = R["indel_off"][i]
        for k in range(R["n_ins"][i] + R["n_del"][i]):
            ipos, ilen, rs, fs = ind[o + k]
            lt_flank, rt_flank = read_seq[:rs], read_seq[rs:]
            lt_ref, rt_ref = ref_seq[:fs], ref_seq[fs:]
            lt_qual, rt_qual = read_qual[:rs], read_qual[rs:]
            padding_base = reference.fetch(chrom, ipos - 1, ipos) if spliced_cigar or not lt_ref else lt_ref[-1]
            if k < R["n_ins"][i]:
                indel_seq = rt_flank[:ilen]
                rt_flank, rt_qual = rt_flank[ilen:], rt_qual[ilen:]
                var = make_var(chrom, ipos, padding_base, padding_base + indel_seq, reference, skip_validation=True)
                read_dict["I"].append((ipos, lt_flank, indel_seq, rt_flank, lt_ref, rt_ref, lt_qual, rt_qual, var))
            else:
                indel_seq = rt_ref[:ilen]
                rt_ref = rt_ref[ilen:]
                var = make_var(chrom, ipos, padding_base + indel_seq, padding_base, reference, skip_validation=True)
                read_dict["D"].append((ipos, lt_flank, indel_seq, rt_flank, lt_ref, rt_ref, lt_qual, rt_qual, var))
        # parse_spliced_read (pileup.pyx:391-436): the integers come from C, the pattern strings are formatted here
        ns, so = R["n_subreads"][i], R["subread_off"][i]
        read_dict["is_covering"] = bool(R["is_covering"][i])
        read_dict["covering_subread"] = (R["covering_start"][i], R["covering_end"][i]) if R["is_covering"][i] else None
        if ns > 1:
            p = R["splice_pos"][i]
            lt_ptrn, rt_ptrn = [], []
            for k in range(ns - 1):
                start, end = sub[so + k][1] + 1, sub[so + k + 1][0] - 1
                if end < p:
                    lt_ptrn.append(f"{start}-{end}")
                elif p < start - 1:
                    rt_ptrn.append(f"{start}-{end}")
            read_dict["is_spliced"] = True
            read_dict["splice_pattern"] = (":".join(lt_ptrn), ":".join(rt_ptrn))
        else:
            read_dict["is_spliced"] = False
            read_dict["splice_pattern"] = ("", "")
        read_dict["intron_pattern"] = (R["intron_start"][i], R["intron_end"][i])
        pileup.append(read_dict)
    pileup = [read for read in pileup if not is_within_intron(read, pos, window)]
    return pileup, sample_factor


class PileupBatch:
    """a locus' reads without per-read Python objects: `batch` (bamio.ReadBatch), `keep` (the records that survive
    fetch_reads, down-sampling and the intron filter), `columns` (bamio.PileupColumns) and `read_table()` for the aligner"""

    def __init__(self, batch, keep, columns, sample_factor):
        self.batch, self.keep, self.columns, self.sample_factor = batch, keep, columns, sample_factor

    def read_table(self):
        """-> (table, off, len, index): SWB_SEQ_PACKED4 read table of the WHOLE region batch and the kept records' indices
        (pair_read entries for swb_align_batch): the bases never exist as Python strings"""
        table, off, length = self.batch.pack4()
        return table, off, length, self.keep


def make_pileup_batch(target, bam, unspl_loc_ref, exclude_duplicates, window, downsamplethresh, basequalthresh) -> PileupBatch:
    """the columnar ingest: fetch_reads + down-sampling + dictize_read's integer core + is_within_intron, no dicts"""
    batch, keep, sample_factor, ref_len, rpos = _select(target, bam, window, downsamplethresh, exclude_duplicates)
    cols = _columns(batch, keep, target, rpos, unspl_loc_ref, basequalthresh)
    r = cols.reads[keep]
    pos = target.pos
    s, e = r["intron_start"], r["intron_end"]
    within = ((s != 0) | (e != 0)) & (s < pos - window) & (pos + window < e)
    return PileupBatch(batch, keep[~within], cols, sample_factor)
