"""BASELINE.json configs[1] at its FULL size (1 M pairs, 150 bp x 400 bp, score + coordinates + CIGAR) on the GPU, checked
through properties that do not need the oracle on every pair:

  * a random sample of the batch against the oracle, bit for bit (the large-batch machinery -- streamed pieces, device job
    lists of hundreds of thousands of entries, re-queue rounds -- must not change a single record);
  * permutation invariance: the same pairs in a shuffled order give, pair by pair, the same records and the same CIGAR words
    (every result depends on its own pair only; this compares all 1 M records and every CIGAR op of two independent runs);
  * internal consistency of every record: coordinates inside the sequences, begin <= end, and the CIGAR's read / reference
    spans equal to the reported intervals (ssw.c:897-900 hands banded_sw exactly those sub-sequences);
  * determinism: a repeated run returns identical bytes.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import swbtest as T  # noqa: E402

pytestmark = pytest.mark.gpu

N = 1_000_000


def _cigar_words(res, arena, order):
    """all CIGAR words of the pairs `order`, concatenated in that order, plus the per-pair lengths"""
    ln = res["cigar_len"][order].astype(np.int64)
    off = res["cigar_off"][order].astype(np.int64)
    tot = int(ln.sum())
    rep = np.repeat(np.arange(order.shape[0]), ln)
    within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
    return arena[off[rep] + within], ln, rep


def test_gpu_full_size_config2_properties():
    from gpuutil import gpu_align

    b = T.make_pairs_fast(N, 150, 400, seed=20261019)
    r1, a1, tm = gpu_align(b)
    assert tm["n_fast"] > 0.99 * N, tm
    assert int((r1["status"] != 0).sum()) == 0 and int((r1["flag"] != 0).sum()) == 0

    # ---- sample vs oracle, bit-exact
    rng = np.random.default_rng(7)
    sub = np.sort(rng.choice(N, 16000, replace=False))
    ro, ao = T.oracle_parallel(b.subset(sub), threads=min(16, os.cpu_count() or 1))
    T.compare(r1[sub].copy(), a1, ro, ao, what="1 M-pair batch, random sample vs oracle")

    # ---- internal consistency of all records
    assert (r1["score1"] <= 450).all() and (r1["score1"] > 0).all()
    for lo, hi, lim in (("ref_begin1", "ref_end1", 400), ("read_begin1", "read_end1", 150)):
        assert (r1[lo] >= 0).all() and (r1[lo] <= r1[hi]).all() and (r1[hi] < lim).all(), (lo, hi)
    allp = np.arange(N)
    words, ln, rep = _cigar_words(r1, a1, allp)
    assert (ln > 0).all()
    op, n = words & 15, (words >> 4).astype(np.int64)
    assert (op <= 2).all() and (n > 0).all()                               # M / I / D only (ssw.c:753-762)
    read_span = np.bincount(rep, weights=np.where(op != 2, n, 0), minlength=N).astype(np.int64)
    ref_span = np.bincount(rep, weights=np.where(op != 1, n, 0), minlength=N).astype(np.int64)
    assert np.array_equal(read_span, r1["read_end1"].astype(np.int64) - r1["read_begin1"] + 1)
    assert np.array_equal(ref_span, r1["ref_end1"].astype(np.int64) - r1["ref_begin1"] + 1)
    # the workload plants one event of 1..10 bp in two thirds of the pairs: gapped CIGARs are common, none absurd
    gapped = np.bincount(rep, weights=(op != 0), minlength=N) > 0
    assert 0.5 < gapped.mean() < 0.75, gapped.mean()

    # ---- permutation invariance over all pairs (second, independent run in shuffled order)
    perm = rng.permutation(N)
    r2, a2, _ = gpu_align(b.subset(perm))
    for f in T.RESULT_FIELDS:
        assert np.array_equal(r2[f], r1[f][perm]), f
    w1, l1, _ = _cigar_words(r1, a1, perm)
    w2, l2, _ = _cigar_words(r2, a2, allp)
    assert np.array_equal(l1, l2) and np.array_equal(w1, w2)

    # ---- determinism
    r3, a3, _ = gpu_align(b)
    for f in T.RESULT_FIELDS:
        assert np.array_equal(r3[f], r1[f]), f
    w3, l3, _ = _cigar_words(r3, a3, allp)
    assert np.array_equal(l3, ln) and np.array_equal(w3, words)
