import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import swbtest as T
from gpuutil import gpu_align
b = T.make_pairs_fast(600000, 100, 260, seed=9, reads_per_window=40)
b.gap_open[::7] = 5
b.gap_ext[::5] = 0
s2 = b.subset(np.array([74347]))
rg, ag, tm = gpu_align(s2)
print(rg)
print(tm)
o2, a2 = T.oracle().align_batch(s2)
print("oracle", o2)
