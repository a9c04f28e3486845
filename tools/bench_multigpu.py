"""Strong scaling through the product's own sharder: ONE process, `MultiGpuAligner` over 1 / 2 / 4 / 8 devices, a FIXED batch
(BASELINE configs[1] shape with shared windows, the realistic layout), end to end from pinned host arrays (every shard uploads the
table slices its pairs touch, aligns, downloads; results stitched in pair order).  Also the locus-sharded reference pipeline under the
wave scheduler (tools/bench_pipeline.py, one worker process per GPU).

    python tools/bench_multigpu.py [--pairs 4000000] [--pipeline-loci 48]
prints one JSON line; run it on a box with several GPUs (`gpurun --gpus 8`)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4_000_000)
    ap.add_argument("--pipeline-loci", type=int, default=48)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import swbtest as T
    from indelpost_b200 import _lib as L
    from indelpost_b200.batch import pack_table
    from indelpost_b200.sharding import MultiGpuAligner

    lib = L.load()
    ndev = lib.swb_device_count()
    b = T.make_pairs_fast(a.pairs, 150, 400, seed=3, reads_per_window=200)
    reads, roff = pack_table(b.reads, b.read_off, b.read_len, bits=2)
    wins, woff = pack_table(b.windows, b.win_off, b.win_len, bits=2)
    # pinned host arrays, like the batched entry's contract asks for (pageable arrays work, at a lower and synchronous copy rate)
    keep = []

    def pin(x):
        buf = L.PinnedBuffer(max(1, x.nbytes))
        v = buf.view(x.dtype, x.size)
        v[:] = x.reshape(-1)
        keep.append(buf)
        return v

    reads, roff, wins, woff = pin(reads), pin(roff), pin(wins), pin(woff)
    rlen, wlen, pr_, pw_, go_, ge_ = pin(b.read_len), pin(b.win_len), pin(b.pair_read), pin(b.pair_win), pin(b.gap_open), pin(b.gap_ext)
    out = {"workload": "cfg2 shape, 200 reads per window, pinned host arrays with 2-bit packed tables, fixed batch (strong scaling), through MultiGpuAligner in one process", "pairs": a.pairs, "devices_on_box": ndev, "by_gpus": {}}
    ref = None
    for n in (1, 2, 4, 8):
        if n > ndev:
            break
        m = MultiGpuAligner(list(range(n)))
        try:
            kw = dict(mat=b.mat, n=5, score_size=2, flag=1, seq_encoding=L.SWB_SEQ_PACKED2, copy=False)
            args = (reads.view(np.int8), roff, rlen, wins.view(np.int8), woff, wlen, pr_, pw_, go_, ge_)
            for _ in range(2):
                r, ar = m.align(*args, **kw)
            t0 = time.perf_counter()
            for _ in range(a.steps):
                r, ar = m.align(*args, **kw)
            dt = (time.perf_counter() - t0) / a.steps
            # (the result arrays are views of the sharder's pinned buffers: read them before close())
            sig = (int(r["score1"].astype(np.int64).sum()), int(r["ref_begin1"].astype(np.int64).sum()), int(r["cigar_len"].astype(np.int64).sum()))
        finally:
            m.close()
        if ref is None:
            ref = sig
        out["by_gpus"][str(n)] = {"gcups": b.cells() / dt / 1e9, "reads_per_s": a.pairs / dt, "ms": dt * 1e3, "identical_to_1gpu": sig == ref}
    if a.pipeline_loci > 0:
        import bench_pipeline as BP

        pl = {}
        for n in (1, 2, 4, 8):
            if n > ndev:
                break
            pl[str(n)] = BP.measure("cfg3", a.pipeline_loci, workers=n, arms=("wave",), devices=tuple(range(n)))["wave"]
        pl["reference_1_process"] = BP.measure("cfg3", min(a.pipeline_loci, 16), workers=1, arms=("reference",))["reference"]
        out["pipeline_cfg3_wave_locus_sharded"] = pl
    print(json.dumps(out))


if __name__ == "__main__":
    main()
