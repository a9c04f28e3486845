"""Aligner glue through which every indelPost caller reaches the hot path (reference
indelpost/localn.pyx:464-472), plus its batched counterpart."""
from __future__ import annotations

from .sswpy import SSW, align_batch


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
