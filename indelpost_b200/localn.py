"""Aligner glue through which every indelPost caller reaches the hot path (reference
indelpost/localn.pyx:464-472), plus its batched counterpart."""
from __future__ import annotations

import re

import numpy as np

from .sswpy import SSW, align_batch, prefetch_alignments, _aligner

cigar_ptrn = re.compile(r"[0-9]+[MIDNSHPX=]")  # localn.pyx:12
_OPS = "MIDNSHP=X"


def make_aligner(ref_seq, match_score, mismatch_penalty):  # localn.pyx:464-467
    aligner = SSW(match_score=match_score, mismatch_penalty=mismatch_penalty)
    aligner.setReference(ref_seq)
    return aligner


def align(aligner, read_seq, gap_open_penalty, gap_extension_penalty):  # localn.pyx:470-472
    aligner.setRead(read_seq)
    return aligner.align(gap_open=gap_open_penalty, gap_extension=gap_extension_penalty)


def align_many(ref_seqs, read_seqs, pair_read, pair_ref, gap_open_penalty, gap_extension_penalty, match_score, mismatch_penalty, device=0):
    """All (read, window, gap-penalty) combinations of one or many loci in one GPU batch: what
    `align(make_aligner(ref), read, go, ge)` returns for each pair, as a list in pair order."""
    return align_batch(read_seqs, ref_seqs, pair_read, pair_ref, gap_open_penalty, gap_extension_penalty,
                       match_score=match_score, mismatch_penalty=mismatch_penalty, device=device)


def generate_grid(auto_adjust_extension_penalty, gap_open_penalty, gap_extension_penalty, target):
    """The (gap_open, gap_extension) points `grid_search` walks for a target indel (reference
    indelpost/varaln.pyx:1122-1146): the user's pair first when it is not (3, 1); (3, 0) moves ahead of (3, 1) for
    indels of 20 bases or more; a single point when auto-adjustment is off.  `target` needs `.indel_seq` only."""
    if not auto_adjust_extension_penalty:
        return [(gap_open_penalty, gap_extension_penalty)]
    head = [(3, 1), (3, 0)] if len(target.indel_seq) < 20 else [(3, 0), (3, 1)]
    grid = head + [(5, 1), (5, 0), (4, 1), (4, 0)]
    if (gap_open_penalty, gap_extension_penalty) != (3, 1):
        grid.insert(0, (gap_open_penalty, gap_extension_penalty))
    return grid


def prefetch_grid_search(target, read_seqs, ref_seqs, auto_adjust_extension_penalty=True, gap_open_penalty=3, gap_extension_penalty=1,
                         match_score=3, mismatch_penalty=2, with_perfect_match=True, device=0):
    """One GPU batch holding every alignment `grid_search` -> `retarget` -> `update_read_info` (varaln.pyx:1164-1243,
    pileup.pyx:639-648, 849) can ask for at one locus: reads x windows x generate_grid(...), plus the `gap_open = len(read)`
    calls of `is_target_by_ssw` (localn.pyx:253-255) and `gap_open = gap_extension = len(read)` of `is_perfect_match`
    (varaln.pyx:1228-1234).  The unmodified per-call code then finds its results in the prefetched set (sswpy.SSW.align).
    Returns the number of alignments computed."""
    grid = list(dict.fromkeys(generate_grid(auto_adjust_extension_penalty, gap_open_penalty, gap_extension_penalty, target)))
    if with_perfect_match:
        # is_target_by_ssw runs with the gap_extension of the grid point that WON the search (varaln.pyx:421, localn.pyx:255)
        grid += [("len", e) for e in dict.fromkeys([gap_extension_penalty] + [e for _, e in grid])] + [("len", "len")]
    return prefetch_alignments(read_seqs, ref_seqs, grid=tuple(grid), match_score=match_score, mismatch_penalty=mismatch_penalty, device=device)


def _pack_cigar(cigarstring):
    """'37M1D113M' -> BAM-packed uint32 ops (len << 4 | op, ssw.h:171-190)"""
    toks = cigar_ptrn.findall(cigarstring or "")
    return np.array([(int(t[:-1]) << 4) | _OPS.index(t[-1]) for t in toks], dtype=np.uint32)


def findall_indels_many(alignments, genome_aln_positions, ref_seqs, read_seqs, report_snvs=False, basequals=None, device=0):
    """findall_indels (localn.pyx:542-621) for many alignments at once: the CIGAR walk with indelPost's
    make_insertion_first reordering (utilities.pyx:360-401) runs on the GPU (k_indels), the dicts -- same keys and
    values as the reference's -- are sliced out of the sequences at the indices it returns."""
    n = len(alignments)
    packed = [_pack_cigar(a.CIGAR) for a in alignments]
    clen = np.array([p.shape[0] for p in packed], dtype=np.int32)
    coff = np.zeros(n, dtype=np.int64)
    if n > 1:
        coff[1:] = np.cumsum(clen[:-1], dtype=np.int64)
    arena = np.concatenate(packed) if n and int(clen.sum()) else np.zeros(0, dtype=np.uint32)
    rs = np.array([a.reference_start for a in alignments], dtype=np.int32)
    qs = np.array([a.read_start for a in alignments], dtype=np.int32)
    off, cnt, rend, recs = _aligner(device).indels_from_cigars(arena, coff, clen, rs, qs)
    out = []
    for k in range(n):
        ref_seq, read_seq, gpos = ref_seqs[k], read_seqs[k], genome_aln_positions[k]
        quals = basequals[k] if basequals is not None else None
        lt_clipped = read_seq[: alignments[k].read_start]
        rt_clipped = read_seq[int(rend[k]):]
        indels = []
        for r in recs[int(off[k]): int(off[k]) + int(cnt[k])]:
            ln, op = int(r["cigar_op"]) >> 4, int(r["cigar_op"]) & 15
            ref_idx, read_idx = int(r["ref_idx"]), int(r["read_idx"])
            d = {"pos": gpos + int(r["pos_off"]), "lt_ref": ref_seq[:ref_idx], "lt_flank": read_seq[:read_idx]}
            if quals:
                d["lt_qual"] = quals[:read_idx]
            if op == 1:
                d.update(indel_type="I", indel_seq=read_seq[read_idx: read_idx + ln], rt_ref=ref_seq[ref_idx:], rt_flank=read_seq[read_idx + ln:],
                         ref_idx=ref_idx, read_idx=read_idx)
                if quals:
                    d["rt_qual"] = quals[read_idx + ln:]
            else:
                d.update(indel_type="D", indel_seq="", del_seq=ref_seq[ref_idx: ref_idx + ln], rt_ref=ref_seq[ref_idx + ln:], rt_flank=read_seq[read_idx:],
                         ref_idx=ref_idx, read_idx=read_idx)
                if quals:
                    d["rt_qual"] = quals[read_idx:]
            d["lt_clipped"] = lt_clipped
            d["rt_clipped"] = rt_clipped
            indels.append(d)
        if not report_snvs:
            out.append(indels)
            continue
        # localn.pyx:594-606: mismatches inside the M tokens, walked between the events the device reported
        snvs = []
        pos = gpos - 1
        ref_idx, read_idx = alignments[k].reference_start, alignments[k].read_start
        events = [(int(r["ref_idx"]), int(r["read_idx"]), int(r["cigar_op"]) >> 4, int(r["cigar_op"]) & 15) for r in recs[int(off[k]): int(off[k]) + int(cnt[k])]]
        events.append((None, int(rend[k]), 0, 0))
        for e_ref, e_read, ln, op in events:
            m = e_read - read_idx                            # matched run up to the next event (or the end of the alignment)
            for i in range(m):
                a, b = ref_seq[ref_idx + i: ref_idx + i + 1], read_seq[read_idx + i: read_idx + i + 1]
                if a != b:
                    snvs.append({"pos": pos + i + 1, "ref": a, "alt": b})
            ref_idx += m; read_idx += m; pos += m
            if op == 1:
                read_idx += ln
            elif op == 2:
                ref_idx += ln; pos += ln
        out.append((indels, snvs))
    return out


def findall_indels(ref_aln, genome_aln_pos, ref_seq, read_seq, report_snvs=False, basequals=None):  # localn.pyx:542
    return findall_indels_many([ref_aln], [genome_aln_pos], [ref_seq], [read_seq], report_snvs=report_snvs,
                               basequals=None if basequals is None else [basequals])[0]
