"""Loader for tests/golden/pairs_*.npz (written by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

import swbtest as T

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("pairs_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "pairs_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"pairs_{name}.npz"))
    opt = lambda k: z[k] if z[k].shape[0] else None
    b = T.Batch(
        z["reads"], z["read_off"], z["read_len"], z["windows"], z["win_off"], z["win_len"],
        z["pair_read"], z["pair_win"], z["gap_open"], z["gap_ext"],
        opt("ref_beg"), opt("ref_len"), opt("mask_len"),
        z["mat"], int(z["n"]), int(z["score_size"]), int(z["flag"]), int(z["filters"]), int(z["filterd"]), 0,
    )
    res = z["results"].view(T.RESULT_DTYPE) if z["results"].dtype != T.RESULT_DTYPE else z["results"]
    return b, res, z["cigars"]
