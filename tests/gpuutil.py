"""Helpers to push a swbtest.Batch through the product's C ABI (indelpost_b200.BatchAligner)."""
import numpy as np

import swbtest as T

_aligner = None


def aligner():
    global _aligner
    if _aligner is None:
        from indelpost_b200 import BatchAligner

        _aligner = BatchAligner(0)
    return _aligner


def gpu_align(b: T.Batch):
    a = aligner()
    res, arena = a.align(
        b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext,
        ref_beg=b.ref_beg, ref_len=b.ref_len, mask_len=b.mask_len, mat=b.mat, n=b.n, score_size=b.score_size, flag=b.flag,
        filters=b.filters, filterd=b.filterd, seq_encoding=b.seq_encoding,
    )
    return res.view(T.RESULT_DTYPE), arena, a.timing()
