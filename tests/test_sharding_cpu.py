"""CPU tests of the multi-GPU host logic: contiguous cell-balanced sharding and result
stitching, single-process and as a world_size-2 gloo job in which each rank aligns its own shard with the CPU
checker (the oracle stands in for the device here; the GPU path itself is covered by -m gpu tests)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import swbtest as T
from indelpost_b200 import sharding as S


def _arrs(b):
    return dict(pair_read=b.pair_read, pair_win=b.pair_win, gap_open=b.gap_open, gap_ext=b.gap_ext,
                ref_beg=b.ref_beg, ref_len=b.ref_len, mask_len=b.mask_len)


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(1)
    cells = rng.integers(1000, 100000, size=10007).astype(np.int64)
    for n in (1, 2, 3, 4, 8):
        bounds = S.shard_bounds(cells, n)
        assert bounds[0][0] == 0 and bounds[-1][1] == cells.shape[0]
        assert all(bounds[k][1] == bounds[k + 1][0] for k in range(n - 1))
        sums = np.array([cells[a:b].sum() for a, b in bounds], dtype=np.float64)
        assert sums.max() <= cells.sum() / n + cells.max()
    assert S.shard_bounds(np.zeros(0, np.int64), 4) == [(0, 0)] * 4
    assert S.shard_bounds(np.array([5], np.int64), 3)[-1] == (1, 1) or sum(b - a for a, b in S.shard_bounds(np.array([5], np.int64), 3)) == 1


def test_sharded_alignment_stitches_to_unsharded_result():
    b = T.make_pairs(900, (40, 150), (100, 400), seed=71, reads_per_window=30, grid=True)
    full_r, full_a = T.oracle().align_batch(b)
    cells = S.pair_cells(b.read_len, b.win_len, b.pair_read, b.pair_win)
    for n in (2, 3, 5):
        parts = []
        for p0, p1 in S.shard_bounds(cells, n):
            r, a = T.oracle().align_batch(b.subset(np.arange(p0, p1)))
            parts.append((r, a))
        r, a = S.stitch(parts)
        T.compare(r.view(T.RESULT_DTYPE), a, full_r, full_a, what=f"{n} shards")


def test_two_rank_gloo_job_matches_single_process(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent('''
        import os, sys, pickle
        sys.path.insert(0, os.environ["SWB_TESTS"]); sys.path.insert(0, os.environ["SWB_ROOT"])
        import numpy as np, torch, torch.distributed as dist
        import swbtest as T
        from indelpost_b200 import sharding as S
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        b = T.make_pairs(600, (40, 150), (100, 400), seed=72, reads_per_window=25, grid=True)   # same seed on every rank
        cells = S.pair_cells(b.read_len, b.win_len, b.pair_read, b.pair_win)
        p0, p1 = S.shard_bounds(cells, world)[rank]
        r, a = T.oracle().align_batch(b.subset(np.arange(p0, p1)))      # this rank's shard only; no data-path collective
        out = [None] * world
        dist.all_gather_object(out, (r, a))                              # test-only gather to compare on rank 0
        if rank == 0:
            rr, aa = S.stitch(out)
            fr, fa = T.oracle().align_batch(b)
            T.compare(rr.view(T.RESULT_DTYPE), aa, fr, fa, what="2-rank gloo")
            print("GLOO_OK", p0, p1)
        dist.barrier()
        dist.destroy_process_group()
    '''))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SWB_TESTS=os.path.join(root, "tests"), SWB_ROOT=root, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29531", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_slice_table_rebases_offsets_and_keeps_bad_indices_out():
    """a shard uploads only the table entries its pairs refer to (sharding.slice_table): index range, rebased byte offsets,
    packed tables by their packed byte extents"""
    from indelpost_b200 import _lib as L
    from indelpost_b200.sharding import slice_table

    blob = np.arange(100, dtype=np.int8)
    off = np.array([0, 10, 30, 60, 90], dtype=np.int64)
    ln = np.array([10, 20, 30, 30, 10], dtype=np.int32)
    b, o, l, i0 = slice_table(blob, off, ln, np.array([3, 2, 2, -1, 99]))
    assert i0 == 2 and l.tolist() == [30, 30] and o.tolist() == [0, 30] and b.tolist() == list(range(30, 90))
    b, o, l, i0 = slice_table(blob, off, ln, np.array([-5, 77]))
    assert b.shape[0] == 0 and o.shape[0] == 0 and i0 == 0
    # 2-bit packed: an entry of len bases occupies ceil(len / 4) bytes
    poff = np.array([0, 3, 8], dtype=np.int64)
    pln = np.array([10, 20, 7], dtype=np.int32)
    b, o, l, i0 = slice_table(np.arange(10, dtype=np.int8), poff, pln, np.array([1, 2]), L.SWB_SEQ_PACKED2)
    assert i0 == 1 and o.tolist() == [0, 5] and b.tolist() == list(range(3, 10))


def test_sampled_bounds_cover_and_balance():
    """sharding.sampled_bounds: contiguous cover of all pairs, balanced by cells within the sampling stride"""
    from indelpost_b200.sharding import pair_cells, sampled_bounds

    rng = np.random.default_rng(5)
    n = 50000
    read_len = rng.integers(50, 250, size=3000).astype(np.int32)
    win_len = rng.integers(200, 1000, size=400).astype(np.int32)
    pr = rng.integers(0, 3000, size=n).astype(np.int32)
    pw = np.sort(rng.integers(0, 400, size=n)).astype(np.int32)
    for k in (1, 2, 3, 8):
        b = sampled_bounds(read_len, win_len, pr, pw, None, None, k)
        assert len(b) == k and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(k - 1))
        cells = pair_cells(read_len, win_len, pr, pw)
        tot = [int(cells[p0:p1].sum()) for p0, p1 in b]
        assert max(tot) <= 1.1 * (sum(tot) / k) + 1
    assert sampled_bounds(read_len, win_len, pr[:0], pw[:0], None, None, 4) == [(0, 0)] * 4
