"""Pipeline-level parity: the UNMODIFIED reference pipeline (oracle/_ref_pipeline: indelpost's compiled modules + the stub
pysam, built by oracle/build_ref_pipeline.py) run twice per locus -- once on its own sswpy/ssw.c, once with the Smith-Waterman
class swapped (`indelpost.localn.SSW`, the name every caller builds its aligners through, localn.pyx:464-467):

  CPU : OracleSSW (the oracle restatement behind the sswpy API)       -> pins the oracle at pipeline level
  GPU : indelpost_b200.SSW (per-call), + prefetch, + the wave scheduler -> the product

and compared: count_alleles (all flag combinations), phase(), the target indel, the contig, and the complete Smith-Waterman
call stream (arguments and the 7-tuple each call returned).  Loci: tests/loci.py::parity_specs() -- all six call sites of
SURVEY.md §3.2 (grid/retarget incl. the window/3 recursion, update_read_info, overhang filter, is_target_by_ssw,
is_perfect_match, decompose_complex_variant), 75-250 bp reads, windows 6-1002 bp."""
from collections import Counter

import numpy as np
import pytest

import loci
import refpipe
import swbtest as T

pytestmark = pytest.mark.skipif(not refpipe.available(), reason="oracle/_ref_pipeline not built (python oracle/build_ref_pipeline.py)")

_OPS = "MIDNSHP=X"


class OracleSSW:
    """the sswpy API (sswpy.pyx:99-304) over the CPU oracle -- test infrastructure"""

    def __init__(self, match_score=2, mismatch_penalty=2):
        self.mat = T.dna_matrix(match_score, mismatch_penalty)
        self.ref = self.read = None

    def setReference(self, reference):
        self.ref = T.encode_dna(reference)

    def setRead(self, read):
        self.read = T.encode_dna(read)

    def align(self, gap_open=3, gap_extension=1, start_idx=0, end_idx=0):
        from indelpost_b200.sswpy import Alignment

        e = len(self.ref) if end_idx == 0 else end_idx
        b = T.batch_from_lists([self.read], [self.ref], [0], [0], gap_open & 0xFF, gap_extension & 0xFF, ref_beg=[start_idx], ref_len=[e - start_idx])
        b.mat = self.mat
        res, arena = T.oracle().align_batch(b)
        r = res[0]
        cig = T.cigar_string(arena, int(r["cigar_off"]), int(r["cigar_len"]))
        return Alignment(cig, int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"]))


_ref_runs = {}


def reference_run(spec):
    """summary + call stream of the reference on its own ssw.c (cached per spec)"""
    key = tuple(sorted(spec.items()))
    if key not in _ref_runs:
        lc = loci.make_locus(**spec)
        calls = []
        out = refpipe.run_locus(lc, calls=calls)
        _ref_runs[key] = (lc, out, calls)
    return _ref_runs[key]


def _assert_same(spec, out_ref, calls_ref, out, calls):
    assert out == out_ref, f"{spec}: pipeline outputs differ\n  reference {out_ref}\n  swapped   {out}"
    assert len(calls) == len(calls_ref), f"{spec}: {len(calls)} SW calls vs {len(calls_ref)} in the reference run"
    for k, (a, b) in enumerate(zip(calls_ref, calls)):
        assert a == b, f"{spec}: SW call {k} differs\n  reference {a[2:]}\n  swapped   {b[2:]}"


def test_specs_cover_every_call_site():
    seen = Counter()
    wins = set()
    for spec in loci.parity_specs():
        lc, out, calls = reference_run(spec)
        lc["_read_set"] = {r["query_sequence"] for r in lc["reads"]}
        for c in calls:
            seen[refpipe.classify_call(c, lc)] += 1
            wins.add(len(c[0]))
    assert len(loci.parity_specs()) >= 50
    for site in ("grid_or_localn.ref", "is_target_by_ssw.mut", "overhang.genome", "overhang.junction", "decompose_complex_variant", "is_perfect_match"):
        assert seen[site] > 0, (site, seen)
    assert {96, 30, 6} <= wins          # window/3 recursion of retarget (pileup.pyx:715-732)
    assert 1002 in wins                 # 250-bp reads, window=167
    assert 199 in wins                  # spliced windows (utilities.pyx:528-575)


def test_oracle_behind_the_reference_pipeline():
    """the oracle restatement, plugged into the unmodified pipeline, changes nothing (every 4th locus: the scalar oracle is slow)"""
    for spec in loci.parity_specs()[::4]:
        lc, out_ref, calls_ref = reference_run(spec)
        calls = []
        out = refpipe.run_locus(loci.make_locus(**spec), ssw_cls=OracleSSW, calls=calls)
        _assert_same(spec, out_ref, calls_ref, out, calls)


@pytest.mark.gpu
def test_gpu_ssw_behind_the_reference_pipeline():
    """zero-change drop-in: every align() is one GPU call"""
    from indelpost_b200 import SSW

    for spec in loci.parity_specs():
        lc, out_ref, calls_ref = reference_run(spec)
        calls = []
        out = refpipe.run_locus(loci.make_locus(**spec), ssw_cls=SSW, calls=calls)
        _assert_same(spec, out_ref, calls_ref, out, calls)


@pytest.mark.gpu
def test_gpu_prefetch_behind_the_reference_pipeline():
    """one prefetch batch per locus (reads x unspliced window x gap grid), then the unmodified control flow"""
    from indelpost_b200 import SSW, clear_prefetched
    from indelpost_b200.localn import prefetch_grid_search

    class _T:
        indel_seq = "A"

    for spec in loci.parity_specs()[::3]:
        lc, out_ref, calls_ref = reference_run(spec)
        lc2 = loci.make_locus(**spec)
        w = lc2["kwargs"]["window"]
        g, pos = lc2["genome"], lc2["pos"]
        windows = [g[max(0, pos - 3 * w): pos + 3 * w]]
        clear_prefetched()
        prefetch_grid_search(_T(), [r["query_sequence"] for r in lc2["reads"]], windows)
        calls = []
        out = refpipe.run_locus(lc2, ssw_cls=SSW, calls=calls)
        clear_prefetched()
        _assert_same(spec, out_ref, calls_ref, out, calls)


@pytest.mark.gpu
def test_gpu_wave_runner_behind_the_reference_pipeline():
    """all loci advance together as wave tasks: every miss of the unmodified control flow is served by ONE merged GPU batch per
    wave (indelpost_b200/wave.py); outputs and call streams identical to the reference's own run"""
    from indelpost_b200 import SSW, clear_prefetched, wave

    specs = loci.parity_specs()
    want = [reference_run(spec) for spec in specs]
    got_calls = [[] for _ in specs]
    sink = refpipe.ThreadCalls()
    pysam = refpipe.load()[1]
    tee = wave.tee_alignment_file(pysam.AlignmentFile)          # reads are registered as the pileup fetches them

    def run(k):
        sink.start(got_calls[k])
        return refpipe.run_locus(loci.make_locus(**specs[k]), swap=False, bam_cls=tee)

    clear_prefetched()
    runner = wave.WaveRunner(max_inflight=64)
    with refpipe.swapped(refpipe.recording(SSW, sink)):
        outs = runner.map(run, range(len(specs)))
    for k, spec in enumerate(specs):
        _assert_same(spec, want[k][1], want[k][2], outs[k], got_calls[k])
    n_calls = sum(len(c) for c in got_calls)
    assert runner.stats["waves"] < n_calls / 100, runner.stats          # ~26 k calls, a few dozen GPU batches
    clear_prefetched()
