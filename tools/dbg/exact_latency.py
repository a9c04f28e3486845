import sys, os, numpy as np, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import swbtest as T
from indelpost_b200 import BatchAligner
al = BatchAligner(0)
for npairs in (1, 10, 100):
    b = T.make_pairs_fast(npairs, 150, 400, seed=3)
    al.upload(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat, n=5, score_size=2, flag=1)
    for _ in range(3): al.compute()
    tm = al.timing()
    print(npairs, {k: round(v, 3) for k, v in tm.items() if k.startswith("ms_")}, tm["n_exact"], tm["n_fast"])
