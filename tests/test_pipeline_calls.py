"""Replay of the reference pipeline's own Smith-Waterman call stream (tests/golden/pipeline_calls.json.gz: every distinct call
`VariantAlignment` + `count_alleles` + `phase` issued on the 53 synthetic loci of tests/loci.py::parity_specs() -- all six call
sites of SURVEY.md 3.2 -- with the results the reference's sswpy/ssw.c returned; written by tests/golden/make_pipeline_golden.py
from the unmodified reference + a stub pysam).

CPU: the oracle reproduces every recorded result.  GPU: the product's batched and per-call entry points do."""
import gzip
import json
import os
from collections import Counter

import numpy as np
import pytest

import swbtest as T
from golden_io import GOLDEN_DIR

_OPS = "MIDNSHP=X"


def _load():
    with gzip.open(os.path.join(GOLDEN_DIR, "pipeline_calls.json.gz"), "rb") as fh:
        return json.loads(fh.read())


def _expected(calls):
    return [tuple(c["out"]) for c in calls]


def _as_batch(doc, calls):
    seqs = [s.encode() for s in doc["seqs"]]
    reads_idx = sorted({c["read"] for c in calls})
    refs_idx = sorted({c["ref"] for c in calls})
    rmap = {s: i for i, s in enumerate(reads_idx)}
    wmap = {s: i for i, s in enumerate(refs_idx)}
    reads = [T.encode_dna(seqs[i]) for i in reads_idx]
    wins = [T.encode_dna(seqs[i]) for i in refs_idx]
    pr = [rmap[c["read"]] for c in calls]
    pw = [wmap[c["ref"]] for c in calls]
    wl = np.array([len(w) for w in wins])[pw]
    s0 = np.array([c["start_idx"] for c in calls])
    e0 = np.array([c["end_idx"] for c in calls])
    e0 = np.where(e0 == 0, wl, e0)
    b = T.batch_from_lists(reads, wins, pr, pw, [c["go"] & 0xFF for c in calls], [c["ge"] & 0xFF for c in calls], ref_beg=s0, ref_len=e0 - s0)
    return b


def _tuples(res, arena):
    out = []
    for k in range(res.shape[0]):
        r = res[k]
        out.append((T.cigar_string(arena, int(r["cigar_off"]), int(r["cigar_len"])), int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]),
                    int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"])))
    return out


def test_fixture_shape():
    doc = _load()
    assert len(doc["calls"]) > 20000 and len(doc["loci"]) >= 50
    sites = Counter(c["site"] for c in doc["calls"])
    for site in ("grid_or_localn.ref", "is_target_by_ssw.mut", "overhang.genome", "overhang.junction", "decompose_complex_variant", "is_perfect_match"):
        assert sites[site] > 0, site
    assert {6, 30, 96, 199, 200, 300, 1002} <= {len(doc["seqs"][c["ref"]]) for c in doc["calls"]}
    kinds = Counter((c["go"], c["ge"]) for c in doc["calls"])
    assert {(3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)} <= set(kinds)          # the gap-penalty grid of varaln.pyx:1127-1143
    assert any(go >= 100 for go, _ in kinds)                                          # localn's go = len(read) aligner (localn.pyx:255)
    for l in doc["loci"]:
        assert sum(l["summary"]["counts"]) > 0


def test_oracle_reproduces_pipeline_calls():
    doc = _load()
    by_matrix = {}
    for c in doc["calls"]:
        by_matrix.setdefault((c["match"], c["mismatch"]), []).append(c)
    for (m, x), calls in by_matrix.items():
        b = _as_batch(doc, calls)
        b.mat = T.dna_matrix(m, x)
        res, arena = T.oracle().align_batch(b)
        assert _tuples(res, arena) == _expected(calls)


@pytest.mark.gpu
def test_gpu_batch_reproduces_pipeline_calls():
    from gpuutil import gpu_align

    doc = _load()
    by_matrix = {}
    for c in doc["calls"]:
        by_matrix.setdefault((c["match"], c["mismatch"]), []).append(c)
    for (m, x), calls in by_matrix.items():
        b = _as_batch(doc, calls)
        b.mat = T.dna_matrix(m, x)
        res, arena, tm = gpu_align(b)
        assert _tuples(res, arena) == _expected(calls)


@pytest.mark.gpu
def test_gpu_sswpy_layer_reproduces_pipeline_calls():
    """through the drop-in Python layer: align_batch on ASCII strings, and make_aligner/align one call at a time"""
    from indelpost_b200 import align_batch
    from indelpost_b200.localn import align, make_aligner

    doc = _load()
    seqs = doc["seqs"]
    calls = [c for c in doc["calls"] if (c["match"], c["mismatch"]) == (3, 2)]
    outs = align_batch(seqs, seqs, [c["read"] for c in calls], [c["ref"] for c in calls],
                       gap_open=[c["go"] for c in calls], gap_extension=[c["ge"] for c in calls],
                       start_idx=[c["start_idx"] for c in calls], end_idx=[c["end_idx"] for c in calls], match_score=3, mismatch_penalty=2)
    assert [tuple(o) for o in outs] == _expected(calls)
    for c in calls[:: max(1, len(calls) // 60)]:
        if c["start_idx"] or c["end_idx"]:
            continue
        got = align(make_aligner(seqs[c["ref"]], 3, 2), seqs[c["read"]], c["go"], c["ge"])
        assert tuple(got) == tuple(c["out"])


class _Target:
    def __init__(self, indel_seq):
        self.indel_seq = indel_seq


def test_generate_grid_follows_the_reference_order():
    """varaln.pyx:1122-1146, case by case"""
    from indelpost_b200.localn import generate_grid

    short, long_ = _Target("ACG"), _Target("A" * 20)
    assert generate_grid(True, 3, 1, short) == [(3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)]
    assert generate_grid(True, 3, 1, long_) == [(3, 0), (3, 1), (5, 1), (5, 0), (4, 1), (4, 0)]
    assert generate_grid(True, 6, 2, short) == [(6, 2), (3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)]
    assert generate_grid(True, 5, 1, long_) == [(5, 1), (3, 0), (3, 1), (5, 1), (5, 0), (4, 1), (4, 0)]
    assert generate_grid(False, 6, 2, short) == [(6, 2)]
    assert generate_grid(False, 3, 1, long_) == [(3, 1)]


def test_prefetch_grid_covers_every_recorded_call():
    """the penalty pairs prefetch_grid_search() submits for a locus are a superset of what the reference pipeline asked for"""
    from indelpost_b200.localn import generate_grid

    doc = _load()
    grid = set(generate_grid(True, 3, 1, _Target("A")))
    missing = Counter()
    for c in doc["calls"]:
        L = len(doc["seqs"][c["read"]])
        go, ge = c["go"], c["ge"]
        if (go, ge) in grid or (go == L and (ge == L or ge in {e for _, e in grid})):
            continue
        missing[(go, ge)] += 1
    assert not missing, missing
