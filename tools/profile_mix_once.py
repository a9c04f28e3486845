"""one upload + one compute of the indelPost penalty mix (for ncu captures; see tools/bench_grid_mix.py for the timed version)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import swbtest as T
from indelpost_b200 import BatchAligner

al = BatchAligner(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
b = T.make_pairs_fast(n, 150, 300, seed=11, reads_per_window=500)
rng = np.random.default_rng(2)
combos = np.array([(3, 1), (5, 1), (3, 0), (5, 0), (4, 1), (4, 0), (150, 1)], dtype=np.uint8)
pick = rng.choice(7, size=n, p=[0.223, 0.18, 0.1344, 0.1344, 0.1344, 0.1344, 0.0594])
b.gap_open = np.ascontiguousarray(combos[pick, 0]); b.gap_ext = np.ascontiguousarray(combos[pick, 1])
al.upload(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat, n=5, score_size=2, flag=1)
al.compute()
print(al.timing())
