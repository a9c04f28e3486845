"""CPU tests of the implicit batching behind a ZERO-CHANGE `SSW.align()` (indelpost_b200/sswpy.py, "implicit batching"): the
unmodified reference pipeline runs on the product's SSW class with no prefetch line and no wave scheduler; every alignment
must come out of a widened batch (here computed by the CPU oracle standing in for the device, as in test_wave_cpu.py), the
outputs and the complete SW call stream must equal the reference's, and the number of device round trips must be a small
fraction of the number of calls."""
import numpy as np
import pytest

import loci
import refpipe
import swbtest as T
import test_wave_cpu as W
from indelpost_b200 import sswpy

pytestmark = pytest.mark.skipif(not refpipe.available(), reason="oracle/_ref_pipeline not built (python oracle/build_ref_pipeline.py)")


def test_zero_change_pipeline_runs_on_widened_batches(monkeypatch):
    batches = []
    monkeypatch.setattr(sswpy, "align_batch", W._oracle_align_batch(batches))
    sswpy.clear_prefetched()
    specs = [dict(s, n_reads=min(s["n_reads"], 40)) for s in loci.parity_specs()[::6]]
    n_calls = 0
    for spec in specs:
        locus = loci.make_locus(**spec)
        want_calls, got_calls = [], []
        want = refpipe.run_locus(locus, calls=want_calls)
        before = len(batches)
        got = refpipe.run_locus(locus, ssw_cls=W._NoGpuSSW, calls=got_calls)          # _single_pair raises: nothing may reach it
        assert got == want, spec
        assert got_calls == want_calls, spec
        n_calls += len(want_calls)
        assert len(batches) - before <= len(want_calls)
    assert sswpy.auto_stats["batches"] == len(batches)
    # grid widening alone saves the other grid points of every (read, window); the recent-read ring saves whole windows
    assert len(batches) < 0.45 * n_calls, (len(batches), n_calls)
    assert sswpy.auto_stats["hits"] > 0.5 * n_calls
    sswpy.clear_prefetched()
    assert not sswpy._AUTO and not sswpy._RECENT


def test_auto_batch_details(monkeypatch):
    batches = []
    monkeypatch.setattr(sswpy, "align_batch", W._oracle_align_batch(batches))
    sswpy.clear_prefetched()
    rng = np.random.default_rng(3)
    win1 = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    win2 = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    reads = [win1[20 + 3 * k: 170 + 3 * k] for k in range(12)]

    def ref_result(read, win, go, ge, ms=3, mm=2):
        b = T.batch_from_lists([T.encode_dna(read)], [T.encode_dna(win)], [0], [0], np.array([go & 0xFF], np.uint8), np.array([ge & 0xFF], np.uint8))
        b.mat = T.dna_matrix(ms, mm)
        res, arena = T.oracle().align_batch(b)
        return sswpy.AlignmentList(res.view(sswpy.L.RESULT_DTYPE), arena)[0]

    # first window: one batch per read for the grid; the first len(read) question fetches those penalties for every read seen so far
    for r in reads:
        a = W._NoGpuSSW(3, 2); a.setReference(win1); a.setRead(r)
        for go, ge in ((3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)):
            assert a.align(gap_open=go, gap_extension=ge) == ref_result(r, win1, go, ge)
    assert len(batches) == len(reads) and batches[-1] == len(sswpy._AUTO_GRID)
    for k, r in enumerate(reads):
        a = W._NoGpuSSW(3, 2); a.setReference(win1); a.setRead(r)
        for go, ge in ((len(r), 1), (len(r), 0), (len(r), len(r))):
            assert a.align(gap_open=go, gap_extension=ge) == ref_result(r, win1, go, ge)
    assert len(batches) == len(reads) + 1 and batches[-1] == len(reads) * len(sswpy._AUTO_LEN_GRID)
    # a new window: the first call brings every recent read along; a fresh aligner object per call (like make_aligner) still hits
    for r in reads:
        a = W._NoGpuSSW(3, 2); a.setReference(win2); a.setRead(r)
        assert a.align(gap_open=4, gap_extension=1) == ref_result(r, win2, 4, 1)
    assert len(batches) == len(reads) + 2 and batches[-1] == len(reads) * len(sswpy._AUTO_GRID)
    # a penalty pair outside both grids travels alone; another matrix has its own cache; sub-range searches are not batched
    a = W._NoGpuSSW(3, 2); a.setReference(win2); a.setRead(reads[0])
    assert a.align(gap_open=7, gap_extension=2) == ref_result(reads[0], win2, 7, 2)
    assert len(batches) == len(reads) + 3 and batches[-1] == 1
    assert a.align(gap_open=7, gap_extension=2) is a.align(gap_open=7, gap_extension=2) and len(batches) == len(reads) + 3
    b = W._NoGpuSSW(2, 2); b.setReference(win2); b.setRead(reads[0])
    assert b.align(gap_open=3, gap_extension=1) == ref_result(reads[0], win2, 3, 1, 2, 2)
    assert len(batches) == len(reads) + 4
    with pytest.raises(AssertionError, match="per-call GPU path"):
        a.align(gap_open=3, gap_extension=1, start_idx=10, end_idx=200)
    # a failing batch falls back to the single-pair path (which reports the error the reference's way)
    def broken(*args, **kw):
        raise sswpy.L.SwbError("no device")
    monkeypatch.setattr(sswpy, "align_batch", broken)
    c = W._NoGpuSSW(3, 2); c.setReference(win1[:50]); c.setRead("ACGTACGTAA")
    with pytest.raises(AssertionError, match="per-call GPU path"):
        c.align()
    # switched off: straight to the single-pair path
    monkeypatch.setattr(sswpy, "AUTO_BATCH", False)
    with pytest.raises(AssertionError, match="per-call GPU path"):
        a.align(gap_open=3, gap_extension=1)
    sswpy.clear_prefetched()
