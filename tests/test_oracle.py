"""CPU tests of the parity oracle itself (no GPU):
   * oracle/ssw_oracle.c reproduces every committed golden vector (generated from the reference's
     own compiled ssw.c by tests/golden/make_golden.py);
   * when oracle/_ref/libssw_ref.so is present (build container, or prebuilt on the GPU box) a seeded
     fuzz compares the restatement with the real reference on fresh inputs.
"""
import numpy as np
import pytest

import swbtest as T
from golden_io import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(name):
    b, res, cig = load_golden(name)
    ro, ao = T.oracle().align_batch(b)
    T.compare(ro, ao, res, cig, what=f"oracle vs golden[{name}]")


def test_golden_covers_both_modes_and_status_paths():
    seen_byte = seen_word = seen_flag2 = seen_null = False
    for name in golden_names():
        _, res, _ = load_golden(name)
        seen_byte |= bool((res["score1"] < 255).any())
        seen_word |= bool((res["score1"] >= 255).any())
        seen_flag2 |= bool((res["flag"] == 2).any())
        seen_null |= bool((res["status"] != 0).any())
    assert seen_byte and seen_word and seen_flag2 and seen_null


FUZZ = [
    dict(n_pairs=600, read_len=150, win_len=400, seed=101),
    dict(n_pairs=600, read_len=(20, 150), win_len=(60, 400), seed=102, grid=True, n_rate=0.01, junk_tail=0.2, low_complexity=0.1),
    dict(n_pairs=600, read_len=(30, 100), win_len=300, seed=103, grid=True, max_indel=20),
    dict(n_pairs=150, read_len=250, win_len=1000, seed=104, grid=True, max_indel=40),
    dict(n_pairs=600, read_len=(60, 130), win_len=300, seed=105, go=5, ge=0, max_indel=15),
    dict(n_pairs=300, read_len=(40, 200), win_len=(100, 500), seed=106, go=2, ge=2),
]


@pytest.mark.parametrize("cfg", FUZZ, ids=[str(c["seed"]) for c in FUZZ])
def test_oracle_matches_compiled_reference(cfg, quiet_stderr):
    if not T.have_ref():
        pytest.skip("oracle/_ref/libssw_ref.so not built (reference tree absent)")
    b = T.make_pairs(**cfg)
    ro, ao = T.oracle().align_batch(b)
    with quiet_stderr():
        rr, ar = T.reference().align_batch(b)
    T.compare(ro, ao, rr, ar, what=f"oracle vs reference {cfg}")


def test_encode_dna_matches_sswpy_lut():
    s = bytes(range(256))
    got = T.encode_dna(s)
    want = np.full(256, 4, dtype=np.int8)
    for ch, v in zip("ACGTUacgtu", [0, 1, 2, 3, 0, 0, 1, 2, 3, 0]):
        want[ord(ch)] = v
    assert (got == want).all()
