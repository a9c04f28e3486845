"""resident-throughput survey over BASELINE.json's other shapes (not the headline bench)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import swbtest as T
from indelpost_b200 import BatchAligner

al = BatchAligner(0)
shapes = [
    ("cfg2 150x400 distinct windows", dict(n_pairs=400000, read_len=150, win_len=400, seed=1)),
    ("cfg2 150x400 200 reads/window", dict(n_pairs=400000, read_len=150, win_len=400, seed=2, reads_per_window=200)),
    ("100x300", dict(n_pairs=400000, read_len=100, win_len=300, seed=3)),
    ("cfg4 250x1000", dict(n_pairs=100000, read_len=250, win_len=1000, seed=4, max_indel=20)),
    ("cfg5 250x2000", dict(n_pairs=60000, read_len=250, win_len=2000, seed=5, max_indel=10)),
    ("75x300", dict(n_pairs=400000, read_len=75, win_len=300, seed=6, max_indel=5)),
    ("50x300", dict(n_pairs=400000, read_len=50, win_len=300, seed=7, max_indel=3)),
]
for name, kw in shapes:
    b = T.make_pairs_fast(**kw)
    n = al.upload(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat, n=5, score_size=2, flag=1)
    al.compute(); al.compute()
    t0 = time.perf_counter(); K = 3
    for _ in range(K): al.compute()
    dt = (time.perf_counter() - t0) / K
    tm = al.timing()
    print(json.dumps({"shape": name, "pairs": b.n_pairs, "gcups": b.cells() / dt / 1e9, "pairs_per_s": b.n_pairs / dt, "ms": dt * 1e3,
                      "fwd": tm["ms_forward"], "rev": tm["ms_reverse"], "band0": tm["ms_band_round0"], "band_rest": tm["ms_band_rest"], "cert": tm["ms_certify"],
                      "n_fast": tm["n_fast"], "n_exact": tm["n_exact"]}))
