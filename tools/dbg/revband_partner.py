import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import swbtest as T
from gpuutil import gpu_align
b = T.make_pairs_fast(600000, 100, 260, seed=9, reads_per_window=40)
b.gap_open[::7] = 5
b.gap_ext[::5] = 0
P = 74347
cand = np.arange(73000, 76000)
sub = b.subset(cand)
ro, ao = T.oracle().align_batch(sub)
nbad = 0
for x in range(len(cand)):
    if cand[x] == P: continue
    s2 = b.subset(np.array([P, cand[x]]))
    rg, ag, tm = gpu_align(s2)
    o2, a2 = T.oracle().align_batch(s2)
    for k in range(2):
        if rg['ref_begin1'][k] != o2['ref_begin1'][k] or rg['read_begin1'][k] != o2['read_begin1'][k]:
            nbad += 1
            if nbad < 10:
                print("MISMATCH with partner", cand[x], "k", k, "gpu", rg[k], "oracle", o2[k], "go/ge", s2.gap_open, s2.gap_ext)
print("done, nbad", nbad)
