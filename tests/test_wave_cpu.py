"""CPU tests of the wave scheduler's host logic (indelpost_b200/wave.py): the unmodified reference pipeline runs as
cooperative tasks, every SSW.align miss is parked and served by a merged batch.  The batch itself is computed by the CPU
oracle here (it stands in for the device, like in test_sharding_cpu.py; the GPU path is covered by -m gpu tests)."""
import numpy as np
import pytest

import loci
import refpipe
import swbtest as T
from indelpost_b200 import sswpy, wave

pytestmark = pytest.mark.skipif(not refpipe.available(), reason="oracle/_ref_pipeline not built (python oracle/build_ref_pipeline.py)")


class _NoGpuSSW(sswpy.SSW):
    """the product's SSW with the per-call GPU path closed: every answer must come from a wave"""

    def _single_pair(self, gap_open, gap_extension, start_idx, search_length):
        raise AssertionError("the per-call GPU path was taken: a request was not served by a wave")


def _oracle_align_batch(calls_log):
    def align_batch(reads, references, pair_read, pair_ref, gap_open=3, gap_extension=1, start_idx=None, end_idx=None,
                    match_score=2, mismatch_penalty=2, device=0, aligner=None):
        b = T.batch_from_lists([T.encode_dna(r) for r in reads], [T.encode_dna(w) for w in references], list(pair_read), list(pair_ref),
                               (np.asarray(gap_open) & 0xFF).astype(np.uint8), (np.asarray(gap_extension) & 0xFF).astype(np.uint8))
        b.mat = T.dna_matrix(match_score, mismatch_penalty)
        res, arena = T.oracle().align_batch(b)
        calls_log.append(len(pair_read))
        return sswpy.AlignmentList(res.view(sswpy.L.RESULT_DTYPE), arena)
    return align_batch


def test_wave_runner_reproduces_the_reference_pipeline(monkeypatch):
    batches = []
    monkeypatch.setattr(sswpy, "align_batch", _oracle_align_batch(batches))
    sswpy.clear_prefetched()
    specs = [dict(s, n_reads=min(s["n_reads"], 40)) for s in loci.parity_specs()[::5]]
    want = []
    for spec in specs:
        calls = []
        want.append((refpipe.run_locus(loci.make_locus(**spec), calls=calls), calls))
    got_calls = [[] for _ in specs]
    sink = refpipe.ThreadCalls()

    tee = wave.tee_alignment_file(refpipe.load()[1].AlignmentFile)      # half of the loci register their reads through the BAM tee ...

    def run(k):
        sink.start(got_calls[k])
        return refpipe.run_locus(loci.make_locus(**specs[k]), swap=False, bam_cls=tee if k % 2 else None)

    runner = wave.WaveRunner(aligner=object(), max_inflight=4)
    with refpipe.swapped(refpipe.recording(_NoGpuSSW, sink)):
        # ... the other half through the `reads` callback
        outs = runner.map(run, range(len(specs)), reads=lambda k: [] if k % 2 else [r["query_sequence"] for r in loci.make_locus(**specs[k])["reads"]])
    for k, spec in enumerate(specs):
        assert outs[k] == want[k][0], spec
        assert got_calls[k] == want[k][1], spec
    n_calls = sum(len(c) for c in got_calls)
    assert runner.stats["tasks"] == len(specs)
    assert runner.stats["waves"] == len(batches) < n_calls / 20          # thousands of calls, a few dozen batches
    assert runner.stats["requests"] >= runner.stats["waves"]
    assert not sswpy._BLOCKS and not sswpy._BLOCK_FIFO                   # every finished locus dropped its blocks
    sswpy.clear_prefetched()


def test_wave_runner_without_speculation_and_task_errors(monkeypatch):
    batches = []
    monkeypatch.setattr(sswpy, "align_batch", _oracle_align_batch(batches))
    sswpy.clear_prefetched()
    spec = dict(loci.parity_specs()[0], n_reads=12)
    calls = []
    want = refpipe.run_locus(loci.make_locus(**spec), calls=calls)

    def run(k):
        if k == 1:
            raise KeyError("boom")
        return refpipe.run_locus(loci.make_locus(**spec), swap=False)

    with refpipe.swapped(_NoGpuSSW):
        runner = wave.WaveRunner(aligner=object(), speculate=False)
        assert runner.map(run, [0, 2]) == [want, want]
        with pytest.raises(KeyError):
            runner.map(run, [0, 1, 2])
        # a failing batch releases every parked task with the error
        def broken(*a, **k):
            raise RuntimeError("device lost")
        monkeypatch.setattr(sswpy, "align_batch", broken)
        with pytest.raises(RuntimeError, match="device lost"):
            wave.WaveRunner(aligner=object()).map(run, [0, 2])
    assert sswpy._RESOLVER is None
    sswpy.clear_prefetched()
