"""resident / end-to-end throughput on indelPost's own penalty mix (the recorded call stream of tests/golden/pipeline_calls.json:
(3,1) 22 %, (5,1) 18 %, (3,0) (5,0) (4,1) (4,0) 13.4 % each, (len(read) mod 256, 1) 6 %), 500 reads per 300-bp window"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import swbtest as T
from indelpost_b200 import BatchAligner

al = BatchAligner(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
b = T.make_pairs_fast(n, 150, 300, seed=11, reads_per_window=500)
rng = np.random.default_rng(2)
combos = np.array([(3, 1), (5, 1), (3, 0), (5, 0), (4, 1), (4, 0), (150, 1)], dtype=np.uint8)
pick = rng.choice(7, size=n, p=[0.223, 0.18, 0.1344, 0.1344, 0.1344, 0.1344, 0.0594])
b.gap_open = np.ascontiguousarray(combos[pick, 0]); b.gap_ext = np.ascontiguousarray(combos[pick, 1])
args = (b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext)
al.upload(*args, mat=b.mat, n=5, score_size=2, flag=1)
for _ in range(2): al.compute()
t0 = time.perf_counter(); K = 3
for _ in range(K): al.compute()
dt = (time.perf_counter() - t0) / K
tm = al.timing()
for _ in range(2): al.align(*args, mat=b.mat, n=5, score_size=2, flag=1, copy=False)
t0 = time.perf_counter()
for _ in range(K): al.align(*args, mat=b.mat, n=5, score_size=2, flag=1, copy=False)
de = (time.perf_counter() - t0) / K
print(json.dumps({"workload": "indelPost penalty mix, 150 bp x 300 bp, 500 reads per window", "pairs": n, "resident_gcups": b.cells() / dt / 1e9, "resident_pairs_per_s": n / dt,
                  "e2e_gcups": b.cells() / de / 1e9, "e2e_pairs_per_s": n / de, "ms_resident": dt * 1e3, "ms_e2e": de * 1e3,
                  "stage_ms": {k: round(tm[k], 3) for k in ("ms_prepare", "ms_forward", "ms_reverse", "ms_traceback", "ms_band_round0", "ms_band_rest", "ms_certify")},
                  "n_fast": tm["n_fast"], "n_exact": tm["n_exact"]}))
