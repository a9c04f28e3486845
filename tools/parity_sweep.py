"""Large randomized parity sweep (not part of the test suite: run on a GPU box when kernels change).
Compares the CUDA path with the CPU oracle on many shapes / penalties / matrices; prints one line per configuration."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np
import swbtest as T
from gpuutil import gpu_align

threads = min(32, os.cpu_count() or 1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
SEED_SHIFT = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfgs = [
    dict(read_len=150, win_len=400, seed=1001, grid=True),
    dict(read_len=(90, 160), win_len=(200, 500), seed=1002, grid=True, max_indel=24, n_rate=0.004),
    dict(read_len=(100, 151), win_len=300, seed=1003, go=3, ge=1, max_indel=12, junk_tail=0.15),
    dict(read_len=(100, 151), win_len=300, seed=1004, go=5, ge=1, max_indel=30, low_complexity=0.2),
    dict(read_len=250, win_len=1000, seed=1005, grid=True, max_indel=40),
    dict(read_len=(200, 256), win_len=(400, 1200), seed=1006, go=4, ge=1, max_indel=22, sub_rate=0.03),
    dict(read_len=(20, 90), win_len=(60, 300), seed=1007, grid=True, max_indel=8),
    dict(read_len=(100, 150), win_len=(260, 420), seed=1008, go=6, ge=2, match=2, mismatch=3, max_indel=15),
    dict(read_len=(120, 150), win_len=400, seed=1009, go=2, ge=1, match=1, mismatch=1, max_indel=10),
    dict(read_len=150, win_len=400, seed=1010, go=3, ge=1, reads_per_window=50, max_indel=10, sub_rate=0.05),
    # free gap extension + long deletions: wide bands (regular, wider than the matrix, doubled)
    dict(read_len=(60, 150), win_len=(250, 500), seed=1011, go=3, ge=0, max_indel=120),
    dict(read_len=(30, 120), win_len=(200, 512), seed=1012, go=5, ge=0, max_indel=200, junk_tail=0.2, low_complexity=0.1),
    dict(read_len=(100, 300), win_len=(300, 700), seed=1013, go=4, ge=0, max_indel=150, sub_rate=0.03),
]
bad = 0
for cfg in cfgs:
    n = N if not isinstance(cfg["read_len"], int) or cfg["read_len"] < 200 else N // 3
    t0 = time.time()
    cfg = dict(cfg, seed=cfg["seed"] + SEED_SHIFT)
    b = T.make_pairs(n, **cfg)
    ro, ao = T.oracle_parallel(b, threads=threads)
    rg, ag, tm = gpu_align(b)
    try:
        T.compare(rg, ag, ro, ao, what=str(cfg))
        print("OK  ", n, "pairs", "fast", tm["n_fast"], "exact", tm["n_exact"], "%.1fs" % (time.time() - t0), cfg, flush=True)
    except AssertionError as e:
        bad += 1
        print("FAIL", str(e)[:1500], flush=True)
T2 = T.make_window_edge_pairs(min(N, 60000), seed=1100 + SEED_SHIFT)
ro, ao = T.oracle_parallel(T2, threads=threads); rg, ag, tm = gpu_align(T2)
try:
    T.compare(rg, ag, ro, ao, what="window edge"); print("OK   window-edge", N, flush=True)
except AssertionError as e:
    bad += 1; print("FAIL", str(e)[:1500], flush=True)
print("sweep done, failing configurations:", bad)
sys.exit(1 if bad else 0)
