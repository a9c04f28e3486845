import sys, os, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import indelpost_b200 as ip
from indelpost_b200 import sswpy, localn
rng = np.random.default_rng(1)
ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
contig = ref[:150] + "TTG" + ref[150:]
reads = [(contig if k % 2 else ref)[int(s):int(s) + 150] for k, s in enumerate(rng.integers(0, 150, 500))]
grid = sswpy.INDELPOST_GRID + (("len", 1),)
for _ in range(3):
    ip.clear_prefetched(); t0 = time.perf_counter(); n = ip.prefetch_alignments(reads, [ref, contig], grid=grid); dt = time.perf_counter() - t0
print("prefetch", n, "alignments in %.2f ms" % (dt * 1e3))
al = localn.make_aligner(ref, 3, 2)
t0 = time.perf_counter()
for rd in reads:
    for go, ge in sswpy.INDELPOST_GRID: localn.align(al, rd, go, ge)
dt = time.perf_counter() - t0
print("3000 per-call align() served from the prefetched set: %.2f ms (%.1f us per call)" % (dt * 1e3, dt / 3000 * 1e6))
ip.clear_prefetched()
al = localn.make_aligner(ref, 3, 2)
t0 = time.perf_counter()
for rd in reads[:100]:
    localn.align(al, rd, 3, 1)
dt = time.perf_counter() - t0
print("100 per-call align() on the GPU, one pair per call: %.2f ms (%.1f us per call)" % (dt * 1e3, dt / 100 * 1e6))
