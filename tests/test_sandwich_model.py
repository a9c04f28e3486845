"""Pins the claim behind DESIGN.md §9's short-read certificate (tools/research/short_read_sandwich*.py, a CPU model; the kernels do
not use it yet): whenever the model certifies a pair whose 8-bit pass is final, Gotoh's forward and reverse outputs ARE the reference's
(oracle), and pairs on which the 8-bit pass really deviates from Gotoh are never certified."""
import dataclasses
import os
import sys

import numpy as np

import swbtest as T

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "research"))

FIELDS = ("score1", "ref_end1", "read_end1", "ref_begin1", "read_begin1")


def test_certified_pairs_have_the_8bit_outputs_and_real_deviations_are_rejected():
    import short_read_sandwich_adversarial as A

    n = 30000
    b = A.make(n, 11)
    r8, _ = T.oracle_parallel(b, threads=min(8, os.cpu_count() or 1))
    r16, _ = T.oracle_parallel(dataclasses.replace(b, score_size=1), threads=min(8, os.cpu_count() or 1))
    diff = np.zeros(n, dtype=bool)
    for f in FIELDS:
        diff |= r8[f] != r16[f]
    final8 = r8["score1"] < 253
    bites = np.nonzero(diff & final8)[0]
    assert bites.shape[0] > 0                                  # the generator does hit the quirk
    mat = b.mat.reshape(b.n, b.n).astype(np.int64)
    for p in bites[:12]:
        assert not A.certified(b, int(p), mat), f"pair {int(p)} deviates from Gotoh but was certified"
    ok = tot = 0
    for p in np.nonzero(final8 & ~diff)[0][:40]:
        c = A.certified(b, int(p), mat)
        if c is None:
            continue
        tot += 1
        if c:
            ok += 1
            assert A.certified.last == tuple(int(r8[f][p]) for f in FIELDS)
    assert tot > 0 and ok >= 0.7 * tot
