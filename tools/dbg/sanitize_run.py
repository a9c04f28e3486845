"""small end-to-end run for compute-sanitizer: every kernel family on tiny inputs"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import swbtest as T
from gpuutil import gpu_align, aligner
for b in (T.make_pairs(600, (40, 150), (100, 400), seed=42, grid=True, n_rate=0.01, max_indel=22),
          T.make_window_edge_pairs(400, seed=5),
          T.make_pairs(200, 250, 1000, seed=7, max_indel=40),
          T.make_pairs(300, (1, 40), (1, 60), seed=8, grid=True, max_indel=3, win_n_rate=0.05)):
    rg, ag, tm = gpu_align(b)
    ro, ao = T.oracle().align_batch(b)
    T.compare(rg, ag, ro, ao, what="sanitize run")
    n = b.n_pairs
    a = aligner()
    off, cnt, rend, recs = a.indels_from_cigars(ag, rg["cigar_off"], rg["cigar_len"], rg["ref_begin1"], rg["read_begin1"])
    print("ok", n, tm["n_fast"], tm["n_exact"], int(cnt.sum()))
