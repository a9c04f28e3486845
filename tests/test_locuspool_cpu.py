"""CPU tests of the locus-parallel driver (indelpost_b200/locuspool.py): worker processes, each running its share of the loci
as wave tasks, give the results of the serial reference run, in order; errors come back with the worker's traceback.  Like in
test_wave_cpu.py the merged batches are computed by the CPU oracle standing in for the device (the GPU path: tools/bench_pipeline.py
and the -m gpu pipeline tests)."""
import types

import pytest

import loci
import refpipe
from indelpost_b200 import locuspool

pytestmark = pytest.mark.skipif(not refpipe.available(), reason="oracle/_ref_pipeline not built (python oracle/build_ref_pipeline.py)")

SPECS = [dict(s, n_reads=min(s["n_reads"], 30)) for s in loci.parity_specs()[::7]]


def _init(rank, device):
    """runs in every worker: the reference pipeline with the product's SSW swapped in, batches served by the oracle"""
    import test_wave_cpu as W
    from indelpost_b200 import sswpy

    refpipe.load()
    sswpy.align_batch = W._oracle_align_batch([])
    refpipe.load()[2].SSW = W._NoGpuSSW
    return types.SimpleNamespace(aligner=object())


def _run(k):
    if k < 0:
        raise KeyError(f"no such locus {k}")
    return k, refpipe.run_locus(loci.make_locus(**SPECS[k]), swap=False)


def _plain(k):
    return k, refpipe.run_locus(loci.make_locus(**SPECS[k]))


def test_locus_pool_equals_serial_reference():
    want = [(k, refpipe.run_locus(loci.make_locus(**SPECS[k]))) for k in range(len(SPECS))]
    with locuspool.LocusPool(_run, workers=2, devices=(0, 1), init=_init, chunk=2) as pool:
        assert pool.map(range(len(SPECS))) == want
        assert pool.stats["tasks"] == len(SPECS) and pool.stats["waves"] >= 2
        assert pool.map([]) == []
        assert pool.map([3, 1]) == [want[3], want[1]]                 # the pool is reusable, results follow the item order
        with pytest.raises(locuspool.LocusPoolError, match="no such locus"):
            pool.map([0, -1, 2])
        assert pool.map([2]) == [want[2]]                             # ... and survives a failing chunk
    with locuspool.LocusPool(_plain, workers=2, mode="plain") as pool:    # the reference's own arm under the same pool
        assert pool.map(range(4)) == want[:4]


def _dies(k):
    import os
    if k == 2:
        os._exit(7)              # a worker killed mid-item (out of memory, a crashed driver ...)
    return k


def test_locus_pool_reports_a_dead_worker():
    pool = locuspool.LocusPool(_dies, workers=2, mode="plain", chunk=1)
    with pytest.raises(locuspool.LocusPoolError, match="exited without a result"):
        pool.map(range(4))
    pool.close()


def test_locus_pool_start_failure():
    def _bad_init(rank, device):          # not picklable by reference from a spawned worker -> start failure is reported
        raise RuntimeError("x")
    with pytest.raises(Exception):
        locuspool.LocusPool(_run, workers=1, init=_bad_init, start_timeout=30)


def test_pool_on_one_bam_file_matches_plain_arm():
    """tools/bench_pipeline.py --pool: every locus in ONE BAM + FASTA, work items are (chrom, pos, ref, alt); the wave arm (batches by the
    CPU oracle here) and the plain reference arm must give identical outputs"""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_pipeline as BP

    BP.CONFIGS["tiny"] = dict(kinds=[("del", 2), ("ins", 3)], n_reads=40, read_len=150, window=50)
    try:
        out = BP.measure_pool("tiny", n_loci=4, workers=2, cpu_oracle=True)
    finally:
        for k in ("SWB_POOL_CPU_ORACLE", "SWB_POOL_ARM", "SWB_POOL_BAM", "SWB_POOL_FA"):
            os.environ.pop(k, None)
    assert out["identical_outputs"] is True and out["pool_wave"]["waves"] >= 2 and out["bam_bytes"] > 0
