import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import swbtest as T
from gpuutil import gpu_align
b = T.make_pairs_fast(64, 150, 400, seed=42)
rg, ag, tm = gpu_align(b)
ro, ao = T.oracle().align_batch(b)
T.compare(rg, ag, ro, ao, what="mini")
print("ok", tm["n_launches"])
