import sys, os, numpy as np, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import swbtest as T
from indelpost_b200 import BatchAligner
al = BatchAligner(0)
for npairs in (100, 1000, 3000, 10000, 50000):
    b = T.make_pairs_fast(npairs, 150, 400, seed=3, reads_per_window=200)
    args = (b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext)
    for _ in range(3): al.align(*args, mat=b.mat, n=5, score_size=2, flag=1, copy=False)
    t0 = time.perf_counter(); K = 10
    for _ in range(K): al.align(*args, mat=b.mat, n=5, score_size=2, flag=1, copy=False)
    dt = (time.perf_counter() - t0) / K
    tm = al.timing()
    print(npairs, "e2e ms %.3f" % (dt * 1e3), "pairs/s %.0f" % (npairs / dt), {k: round(v, 3) for k, v in tm.items() if k in ("ms_total", "ms_forward", "ms_reverse", "ms_traceback", "ms_band_round0", "ms_band_rest", "ms_certify")}, tm["n_launches"])
