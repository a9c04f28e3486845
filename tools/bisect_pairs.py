"""debug helper: run every pair of a golden set in its own process through the GPU path and report failures"""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
name = sys.argv[1]
if len(sys.argv) > 2:
    import numpy as np, swbtest as T
    from golden_io import load_golden
    from gpuutil import gpu_align
    b, res, cig = load_golden(name)
    idx = [int(x) for x in sys.argv[2].split(",")]
    sb = b.subset(idx)
    rg, ag, tm = gpu_align(sb)
    ro, ao = T.oracle().align_batch(sb)
    try:
        T.compare(rg, ag, ro, ao, what=str(idx))
        print("OK", idx, "fast", tm["n_fast"], "exact", tm["n_exact"])
    except AssertionError as e:
        print("MISMATCH", idx, str(e)[:1500])
    sys.exit(0)
from golden_io import load_golden
b, _, _ = load_golden(name)
for p in range(b.n_pairs):
    r = subprocess.run([sys.executable, __file__, name, str(p)], capture_output=True, text=True)
    out = (r.stdout.strip().splitlines() or ["<no output>"])
    if not out[-1].startswith("OK"):
        rl = int(b.read_len[b.pair_read[p]]); wl = int(b.win_len[b.pair_win[p]])
        print("pair", p, "rl", rl, "wl", wl, "go", int(b.gap_open[p]), "ge", int(b.gap_ext[p]), "rb", None if b.ref_beg is None else int(b.ref_beg[p]), "->", "\n".join(out[-12:])[:1800], r.stderr.strip()[-300:])
print("bisect done")
