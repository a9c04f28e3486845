"""one compute over short reads (44-84 bp, the zone where the 8-bit pass is final and its scores pass 128) + the indel
extraction kernel (for ncu captures)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import swbtest as T
from indelpost_b200 import BatchAligner

al = BatchAligner(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000
b = T.make_pairs_fast(n, 75, 300, seed=6, max_indel=5)
al.upload(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat, n=5, score_size=2, flag=1)
al.compute()
t0 = time.perf_counter(); al.compute(); dt = time.perf_counter() - t0
tm = al.timing()
al.indels(n)
print(json.dumps({"shape": "75x300", "pairs": n, "gcups": b.cells() / dt / 1e9, "ms": dt * 1e3, "n_fast": tm["n_fast"], "n_exact": tm["n_exact"],
                  "fwd": tm["ms_forward"], "rev": tm["ms_reverse"], "tb": tm["ms_traceback"]}))
