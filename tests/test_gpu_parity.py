"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, must
reproduce bit-for-bit (a) every committed golden vector generated from the reference's own ssw.c and
(b) the CPU oracle on fresh seeded inputs, for every s_align field and every CIGAR op."""
import json
import os
import sys

import numpy as np
import pytest

import swbtest as T
from golden_io import GOLDEN_DIR, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_gpu_reproduces_golden(name):
    from gpuutil import gpu_align

    b, res, cig = load_golden(name)
    rg, ag, _ = gpu_align(b)
    T.compare(rg, ag, res, cig, what=f"gpu vs golden[{name}]")


FUZZ = [
    dict(n_pairs=4000, read_len=150, win_len=400, seed=201),
    dict(n_pairs=4000, read_len=150, win_len=400, seed=202, reads_per_window=200, n_rate=0.005),
    dict(n_pairs=4000, read_len=(20, 150), win_len=(60, 400), seed=203, grid=True, n_rate=0.01, junk_tail=0.2, low_complexity=0.1),
    dict(n_pairs=4000, read_len=(30, 100), win_len=300, seed=204, grid=True, max_indel=20),
    dict(n_pairs=600, read_len=250, win_len=1000, seed=205, grid=True, max_indel=40),
    dict(n_pairs=300, read_len=250, win_len=2000, seed=206, max_indel=200),
    dict(n_pairs=3000, read_len=(1, 40), win_len=(1, 60), seed=207, grid=True, max_indel=3, win_n_rate=0.05, n_rate=0.05),
    dict(n_pairs=4000, read_len=(60, 130), win_len=300, seed=208, go=5, ge=0, max_indel=15),
    dict(n_pairs=2000, read_len=(40, 200), win_len=(100, 500), seed=209, go=2, ge=2),
    dict(n_pairs=2000, read_len=(40, 200), win_len=(100, 500), seed=210, go=1, ge=3),
    dict(n_pairs=2000, read_len=(40, 200), win_len=(100, 500), seed=211, go=0, ge=0),
    dict(n_pairs=2000, read_len=(40, 200), win_len=(100, 500), seed=212, go=6, ge=2, match=1, mismatch=4),
    dict(n_pairs=1000, read_len=(200, 600), win_len=(300, 900), seed=213, grid=True, max_indel=60),
    # free gap extension + long deletions: wide bands (swb_bandwarp.cuh), regular and wider than the matrix, doubled ones
    dict(n_pairs=4000, read_len=(60, 150), win_len=(250, 500), seed=214, go=3, ge=0, max_indel=120),
    dict(n_pairs=4000, read_len=(30, 120), win_len=(200, 512), seed=215, go=5, ge=0, max_indel=200, junk_tail=0.2, low_complexity=0.1),
    dict(n_pairs=3000, read_len=(100, 300), win_len=(300, 700), seed=216, go=4, ge=0, max_indel=150, sub_rate=0.03),
]


@pytest.mark.parametrize("cfg", FUZZ, ids=[str(c["seed"]) for c in FUZZ])
def test_gpu_matches_oracle(cfg):
    from gpuutil import gpu_align

    b = T.make_pairs(**cfg)
    ro, ao = T.oracle().align_batch(b)
    rg, ag, _ = gpu_align(b)
    T.compare(rg, ag, ro, ao, what=f"gpu vs oracle {cfg}")


def test_gpu_ascii_encoding_and_bad_input():
    from gpuutil import gpu_align

    rng = np.random.default_rng(3)
    win = "".join(rng.choice(list("ACGT"), 300))
    reads = [win[50:200], win[50:120].lower() + "NRY" + win[123:200], win[10:100].replace("T", "U"), ""]
    b = T.batch_from_lists([r.encode() for r in reads], [win.encode(), b""], [0, 1, 2, 3, 0], [0, 0, 0, 0, 1], 3, 1, seq_encoding=1)
    rg, ag, _ = gpu_align(b)
    ok = b.subset([0, 1, 2])
    ro, ao = T.oracle().align_batch(ok)
    T.compare(rg[:3], ag, ro, ao, what="ascii")
    assert rg["status"][3] == 2 and rg["status"][4] == 2          # empty read / empty window -> SWB_ERR_BAD_INPUT
    assert rg["ref_begin1"][3] == -1 and rg["cigar_len"][3] == 0


def test_gpu_aliasing_table_entries():
    """table entries that share blob bytes: fine for codes (nothing is rewritten), rejected for ASCII input -- the device encodes
    ASCII tables in place and DNA_BASE_LUT is not idempotent ('A' -> 0 -> 4), so a shared byte would be encoded twice"""
    from gpuutil import aligner, gpu_align
    from indelpost_b200 import _lib as L

    rng = np.random.default_rng(12)
    win = rng.integers(0, 4, 400).astype(np.int8)
    blob = np.concatenate([win[40:240], win[300:380]])           # read 0 = blob[0:150], read 1 = blob[50:200] (overlapping), read 2 = blob[200:280]
    b = T.Batch(blob, np.array([0, 50, 200], np.int64), np.array([150, 150, 80], np.int32), win, np.array([0, 0], np.int64), np.array([400, 300], np.int32),
                np.array([0, 1, 2, 1], np.int32), np.array([0, 0, 0, 1], np.int32), np.full(4, 3, np.uint8), np.full(4, 1, np.uint8))
    rg, ag, _ = gpu_align(b)
    flat = T.batch_from_lists([blob[0:150], blob[50:200], blob[200:280]], [win, win[:300]], [0, 1, 2, 1], [0, 0, 0, 1], 3, 1)
    ro, ao = T.oracle().align_batch(flat)
    T.compare(rg, ag, ro, ao, what="aliasing code tables")
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    b.reads = lut[b.reads].view(np.int8)
    b.windows = lut[b.windows].view(np.int8)
    b.seq_encoding = 1
    with pytest.raises(L.SwbError, match="must not overlap"):
        gpu_align(b)
    b.read_off = np.array([200, 0, 50], np.int64)                # disjoint but not ascending
    b.read_len = np.array([80, 40, 150], np.int32)
    with pytest.raises(L.SwbError, match="ascending"):
        gpu_align(b)
    rg2, ag2, _ = gpu_align(flat)                                 # the context is usable after the rejected calls
    T.compare(rg2, ag2, ro, ao, what="after rejected batches")


def test_python_layer_rejects_short_per_pair_arrays():
    from gpuutil import aligner

    b = T.make_pairs(64, (60, 100), 200, seed=3)
    a = aligner()
    with pytest.raises(ValueError, match="one entry per pair"):
        a.align(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open[:10], b.gap_ext, mat=b.mat)
    with pytest.raises(ValueError, match="equal sizes"):
        a.align(b.reads, b.read_off[:-1], b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat)
    from indelpost_b200 import align_batch
    with pytest.raises(ValueError, match="one entry per pair"):
        align_batch(["ACGT" * 10], ["ACGT" * 30], [0, 0, 0], [0, 0, 0], gap_open=[3, 3])


def test_sswpy_api_matches_reference_golden():
    """the reference's own sswpy.SSW outputs (tests/golden/sswpy_api.json) through our SSW class"""
    from indelpost_b200 import SSW, align_batch

    cases = json.load(open(os.path.join(GOLDEN_DIR, "sswpy_api.json")))
    batch_reads, batch_refs, kws, wants = [], [], [], []
    for c in cases:
        ref, read = c["ref"], c["read"]
        if c["bytes_input"]:
            ref, read = ref.encode(), read.encode()
        a = SSW(c["match"], c["mismatch"])
        a.setReference(ref)
        a.setRead(read)
        if c["err"]:
            with pytest.raises(ValueError):
                a.align(**c["kw"])
            continue
        got = a.align(**c["kw"])
        assert list(got) == c["out"], (c["kw"], list(got), c["out"])
        assert got.CIGAR == c["out"][0] and got.optimal_score == c["out"][1]
        if c["match"] == 3:
            batch_reads.append(read); batch_refs.append(ref); kws.append(c["kw"]); wants.append(c["out"])
    n = len(batch_reads)
    outs = align_batch(
        batch_reads, batch_refs, list(range(n)), list(range(n)),
        gap_open=[k.get("gap_open", 3) for k in kws], gap_extension=[k.get("gap_extension", 1) for k in kws],
        start_idx=[k.get("start_idx", 0) for k in kws], end_idx=[k.get("end_idx", 0) for k in kws],
        match_score=3, mismatch_penalty=2,
    )
    assert [list(o) for o in outs] == wants


def test_link_level_dropin_of_reference_sswpy():
    """the reference's UNMODIFIED sswpy.pyx, cythonized against include/compat/ssw.h and linked against libswb200.so
    instead of ssw.c (oracle/build_ref_sswpy_dropin.py), returns the reference's own outputs (tests/golden/sswpy_api.json) -- every
    SSW.align() of that module is a GPU call through ssw_init / ssw_align / align_destroy"""
    d = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "oracle", "_ref_sswpy", "sswpy_dropin")
    if not os.path.isdir(d):
        pytest.skip("oracle/_ref_sswpy not built (needs the reference tree at build time)")
    sys.path.insert(0, os.path.abspath(d))
    try:
        import sswpy as dropin
    finally:
        sys.path.pop(0)
    assert "oracle/_ref_sswpy" in dropin.__file__.replace(os.sep, "/")
    cases = json.load(open(os.path.join(GOLDEN_DIR, "sswpy_api.json")))
    n_ok = 0
    for c in cases:
        ref, read = c["ref"], c["read"]
        if c["bytes_input"]:
            ref, read = ref.encode(), read.encode()
        a = dropin.SSW(c["match"], c["mismatch"])
        a.setReference(ref)
        a.setRead(read)
        if c["err"]:
            with pytest.raises(ValueError):
                a.align(**c["kw"])
            continue
        got = a.align(**c["kw"])
        assert list(got) == c["out"], (c["kw"], list(got), c["out"])
        n_ok += 1
    assert n_ok >= 15
    # the reference's convenience wrappers on top of it (sswpy.pyx:339-396) and ours give the same
    from indelpost_b200 import force_align, format_force_align

    read, ref = "ACGTTGCATGCATTACGATCGATC", "GGGGACGTTGCATGCATTACGATCGATCTTTT"
    want = dropin.force_align(read, ref)
    got = force_align(read, ref)
    assert list(got) == list(want)
    assert format_force_align(read, ref, got) == dropin.format_force_align(read, ref, want)
    with pytest.raises(ValueError, match="No solution found"):
        force_align("A", "CCCCCC")
    with pytest.raises(ValueError, match="one overhang"):
        force_align(read, ref, force_overhang=True)


def test_sswpy_memo_returns_identical_results():
    """repeated SSW.align calls with identical inputs are served from the aligner's memo (update_read_info
    repeats retarget's alignment); a new reference or different penalties must not hit it"""
    from indelpost_b200 import SSW

    rng = np.random.default_rng(4)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    read = ref[60:130] + "GG" + ref[130:200]
    a = SSW(3, 2)
    a.setReference(ref)
    a.setRead(read)
    first = a.align(3, 1)
    assert a.align(3, 1) is first                      # memo hit: the very same tuple
    other = a.align(5, 0)
    assert other is not first
    a.setRead(read[:100])
    shorter = a.align(3, 1)
    assert shorter != first
    a.setRead(read)
    assert a.align(3, 1) == first
    a.setReference(ref[10:])
    moved = a.align(3, 1)
    assert moved.reference_start == first.reference_start - 10 and moved.CIGAR == first.CIGAR


def test_prefetch_serves_the_per_call_api():
    """prefetch_alignments computes reads x windows x gap grid in one batch; the unmodified per-call API (make_aligner /
    align, localn.pyx:464-472) then answers from it with exactly the tuples it would have computed"""
    import indelpost_b200 as ip
    from indelpost_b200 import localn, sswpy

    rng = np.random.default_rng(8)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 320))
    contig = ref[:150] + "TTGCA" + ref[150:]
    reads = []
    for k in range(24):
        s0 = int(rng.integers(0, 150))
        src = contig if k % 2 else ref
        reads.append(src[s0: s0 + 120])
    ip.clear_prefetched()
    direct = {}
    for w, win in enumerate((ref, contig)):
        al = localn.make_aligner(win, 3, 2)
        for r, rd in enumerate(reads[:6]):
            for go, ge in ((3, 1), (4, 0), (len(rd), 1)):
                direct[(w, r, go, ge)] = localn.align(al, rd, go, ge)
    grid = sswpy.INDELPOST_GRID + (("len", 1),)
    n = ip.prefetch_alignments(reads, [ref, contig], grid=grid, match_score=3, mismatch_penalty=2)
    assert n == len(reads) * 2 * len(grid)
    for w, win in enumerate((ref, contig)):
        al = localn.make_aligner(win, 3, 2)                # a fresh aligner: its own memo is empty
        for r, rd in enumerate(reads):
            for go, ge in ((3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0), (len(rd), 1)):
                got = localn.align(al, rd, go, ge)
                hit = sswpy.prefetched(sswpy.dna_score_matrix(3, 2).tobytes(), sswpy.seq_id(rd.encode()), sswpy.seq_id(win.encode()), go & 0xFF, ge & 0xFF, len(rd))
                assert hit is got                              # served from the prefetched set
                if (w, r, go, ge) in direct:
                    assert got == direct[(w, r, go, ge)]
    ip.clear_prefetched()


def test_prefetch_grid_search_covers_the_locus_calls():
    """localn.prefetch_grid_search: reads x windows x generate_grid (+ the len(read) penalties of is_target_by_ssw and
    is_perfect_match) in one batch; every per-call alignment of the locus is then a prefetched tuple equal to a direct call"""
    import indelpost_b200 as ip
    from indelpost_b200 import localn, sswpy

    class Target:
        indel_seq = "TTGCA"

    rng = np.random.default_rng(18)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
    contig = ref[:140] + "TTGCA" + ref[140:]
    reads = [(contig if k % 2 else ref)[s0: s0 + 100] for k, s0 in enumerate(rng.integers(0, 190, 16))]
    ip.clear_prefetched()
    direct = {}
    for w, win in enumerate((ref, contig)):
        al = localn.make_aligner(win, 3, 2)
        for r in (0, 5, 11):
            for go, ge in ((5, 0), (len(reads[r]), 1), (len(reads[r]), len(reads[r]))):
                direct[(w, r, go, ge)] = localn.align(al, reads[r], go, ge)
    n = localn.prefetch_grid_search(Target(), reads, [ref, contig], True, 3, 1, 3, 2)
    assert n == len(reads) * 2 * 9          # 6 grid points + (len, 1) + (len, 0) + (len, len)
    for w, win in enumerate((ref, contig)):
        al = localn.make_aligner(win, 3, 2)
        for r, rd in enumerate(reads):
            for go, ge in localn.generate_grid(True, 3, 1, Target()) + [(len(rd), 1), (len(rd), len(rd))]:
                got = localn.align(al, rd, go, ge)
                hit = sswpy.prefetched(sswpy.dna_score_matrix(3, 2).tobytes(), sswpy.seq_id(rd.encode()), sswpy.seq_id(win.encode()), go & 0xFF, ge & 0xFF, len(rd))
                assert hit is got
                if (w, r, go, ge) in direct:
                    assert got == direct[(w, r, go, ge)]
    ip.clear_prefetched()


def test_c_abi_single_pair_entry_points():
    """ssw_init / ssw_align / align_destroy / init_destroy exactly as sswpy.pyx's extern block binds them"""
    import ctypes as C

    from indelpost_b200 import _lib as L

    lib = L.load()
    b = T.make_pairs(40, (30, 150), 300, seed=31, grid=True)
    ro, ao = T.oracle().align_batch(b)
    mat = np.ascontiguousarray(b.mat, dtype=np.int8)
    for p in range(b.n_pairs):
        rd = np.ascontiguousarray(b.reads[b.read_off[p] : b.read_off[p] + b.read_len[p]])
        w = b.pair_win[p]
        wn = np.ascontiguousarray(b.windows[b.win_off[w] : b.win_off[w] + b.win_len[w]])
        prof = lib.ssw_init(rd.ctypes.data, int(rd.shape[0]), mat.ctypes.data, 5, 2)
        ml = max(15, int(rd.shape[0]) // 2)
        a = lib.ssw_align(prof, wn.ctypes.data, int(wn.shape[0]), int(b.gap_open[p]), int(b.gap_ext[p]), 1, 0, 0, ml)
        assert a
        r = a.contents
        got = (r.score1, r.score2, r.ref_begin1, r.ref_end1, r.read_begin1, r.read_end1, r.ref_end2, r.cigarLen, r.flag)
        want = tuple(int(ro[f][p]) for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "flag"))
        assert got == want, (p, got, want)
        cg = [r.cigar[i] for i in range(r.cigarLen)]
        assert cg == [int(v) for v in ao[int(ro["cigar_off"][p]) : int(ro["cigar_off"][p]) + int(ro["cigar_len"][p])]]
        lib.align_destroy(a)
        lib.init_destroy(prof)


@pytest.mark.parametrize("scalar_lanes", [False, True], ids=["packed", "scalar"])
@pytest.mark.parametrize("cfg", [FUZZ[0], FUZZ[2], FUZZ[3], FUZZ[9], FUZZ[10]], ids=["150x400", "mixed", "short", "go<ge", "go=ge=0"])
def test_gpu_exact_path_alone_matches_oracle(cfg, scalar_lanes, monkeypatch):
    """the exact striped-emulation kernels (packed two-lanes-per-thread variant and the scalar-lane one) must stay
    bit-exact on inputs the DPX fast path normally takes"""
    from gpuutil import gpu_align

    monkeypatch.setenv("SWB200_NO_FAST", "1")
    if scalar_lanes:
        monkeypatch.setenv("SWB200_OPT", "2")
    b = T.make_pairs(**{**cfg, "n_pairs": 1500})
    ro, ao = T.oracle().align_batch(b)
    rg, ag, tm = gpu_align(b)
    assert tm["n_fast"] == 0
    T.compare(rg, ag, ro, ao, what=f"exact-only gpu vs oracle {cfg}")


def test_gpu_fast_path_is_used_for_the_headline_workload():
    from gpuutil import gpu_align

    b = T.make_pairs_fast(20000, 150, 400, seed=5)
    rg, ag, tm = gpu_align(b)
    assert tm["n_fast"] > 0.9 * b.n_pairs, tm
    sub = np.arange(0, b.n_pairs, 10)
    ro, ao = T.oracle().align_batch(b.subset(sub))
    rs = rg[sub].copy()
    # re-base cigar offsets of the subset for the comparison
    T.compare(rs, ag, ro, ao, what="fast path vs oracle (cfg2)")


@pytest.mark.parametrize("seed", [301, 302])
def test_gpu_overflow_decision_stress(seed):
    """the provisional 16-bit results of the fast path must survive ssw.c's 8-bit-first escalation rule on
    inputs built to sit in the zone where the 8-bit pass' signed lazy-F test misbehaves"""
    from gpuutil import gpu_align

    b = T.make_overflow_zone_pairs(24000, seed=seed)
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    rg, ag, tm = gpu_align(b)
    T.compare(rg, ag, ro, ao, what=f"overflow-zone stress seed={seed}")
    assert tm["n_fast"] > 0


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_gpu_reverse_band_window_edges(seed):
    """banded reverse pass (swb_revband.cuh): alignments anchored at a window edge with indels next to it"""
    from gpuutil import gpu_align

    b = T.make_window_edge_pairs(6000, seed=seed)
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    rg, ag, tm = gpu_align(b)
    T.compare(rg, ag, ro, ao, what=f"window-edge reverse stress seed={seed}")
    assert tm["n_fast"] > 0


def test_gpu_pipelined_batch_equals_single_pass(monkeypatch):
    """large batches go through the streamed one-shot path of swb_align_batch (copies cut into pieces, forward sweeps
    queued per piece) or, on request, through the two-lane chunked pipeline (table slices, index rebasing, CIGAR arena
    stitched from chunks): both must equal the single-pass path and the oracle"""
    from gpuutil import gpu_align

    b = T.make_pairs_fast(600000, 100, 260, seed=9, reads_per_window=40)
    # shuffle penalties a bit and add bad indices at chunk edges
    b.gap_open[::7] = 5
    b.gap_ext[::5] = 0
    b.pair_read[12345] = -1
    b.pair_win[250000] = b.n_windows + 3
    assert b.n_pairs >= 524288
    r2, a2, tm2 = gpu_align(b)                                   # default: streamed
    monkeypatch.setenv("SWB200_CHUNK_PAIRS", "70000")          # two-lane pipeline, several chunks on a test-sized batch
    r1, a1, tm1 = gpu_align(b)
    monkeypatch.delenv("SWB200_CHUNK_PAIRS")
    monkeypatch.setenv("SWB200_NO_PIPELINE", "1")
    r0, a0, tm0 = gpu_align(b)
    T.compare(r1, a1, r0, a0, what="two-lane pipeline vs single pass")
    T.compare(r2, a2, r0, a0, what="streamed vs single pass")
    assert r1["status"][12345] == 2 and r1["status"][250000] == 2
    assert r2["status"][12345] == 2 and r2["status"][250000] == 2
    sub = np.arange(0, b.n_pairs, 97)
    sub = sub[(sub != 12345) & (sub != 250000)]
    ro, ao = T.oracle_parallel(b.subset(sub), threads=min(16, os.cpu_count() or 1))
    T.compare(r2[sub].copy(), a2, ro, ao, what="streamed vs oracle")


def test_gpu_parts_equal_single_part(monkeypatch):
    """SWB200_PARTS=k cuts a batch into pair ranges with their own job lists (a part's reverse / traceback stages run beside the next
    part's forward sweep): same results as one part, in the resident path and in the streamed one-shot path (pieces cut at the
    part boundaries, 2-bit packed tables)"""
    from gpuutil import gpu_align

    b = T.make_pairs(9001, (40, 150), (150, 400), seed=91, grid=True, n_rate=0.004, max_indel=30)
    r0, a0, _ = gpu_align(b)
    monkeypatch.setenv("SWB200_PARTS", "3")
    r1, a1, _ = gpu_align(b)
    T.compare(r1, a1, r0, a0, what="3 parts vs 1 part (single pass)")
    big = T.make_pairs_fast(300001, 120, 300, seed=14, reads_per_window=30)
    monkeypatch.setenv("SWB200_PARTS", "4")
    r2, a2, _ = gpu_align(_packed(big, 2))
    monkeypatch.delenv("SWB200_PARTS")
    monkeypatch.setenv("SWB200_NO_PIPELINE", "1")
    r3, a3, _ = gpu_align(big)
    T.compare(r2, a2, r3, a3, what="4 parts streamed vs 1 part single pass")


def test_gpu_streamed_batch_unordered_tables(monkeypatch):
    """streamed path with pairs that refer to the sequence tables in random order (the upload frontier jumps to the
    end with the first piece) and ASCII input"""
    from gpuutil import gpu_align

    b = T.make_pairs_fast(300000, 80, 200, seed=11, reads_per_window=25)
    rng = np.random.default_rng(5)
    perm = rng.permutation(b.n_pairs)
    b = b.subset(perm)
    r2, a2, _ = gpu_align(b)
    monkeypatch.setenv("SWB200_NO_PIPELINE", "1")
    r0, a0, _ = gpu_align(b)
    T.compare(r2, a2, r0, a0, what="streamed (shuffled pairs) vs single pass")
    sub = np.arange(0, b.n_pairs, 211)
    ro, ao = T.oracle_parallel(b.subset(sub), threads=min(16, os.cpu_count() or 1))
    T.compare(r2[sub].copy(), a2, ro, ao, what="streamed (shuffled pairs) vs oracle")


def test_gpu_streamed_batch_ascii_subranges_and_masks(monkeypatch):
    """streamed path with every optional input: ASCII sequences (encoded piece by piece on the device), per-pair
    ref_beg / ref_len sub-ranges and explicit mask lengths"""
    from gpuutil import gpu_align

    b = T.make_pairs_fast(280000, 90, 240, seed=21, reads_per_window=30)
    rng = np.random.default_rng(3)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    b.reads = lut[b.reads].view(np.int8)
    b.windows = lut[b.windows].view(np.int8)
    b.seq_encoding = 1
    b.ref_beg = rng.integers(0, 30, size=b.n_pairs).astype(np.int32)
    b.ref_len = (240 - b.ref_beg - rng.integers(0, 30, size=b.n_pairs)).astype(np.int32)
    b.mask_len = rng.integers(10, 60, size=b.n_pairs).astype(np.int32)
    r2, a2, _ = gpu_align(b)
    monkeypatch.setenv("SWB200_NO_PIPELINE", "1")
    r0, a0, _ = gpu_align(b)
    T.compare(r2, a2, r0, a0, what="streamed (ASCII, sub-ranges, masks) vs single pass")
    sub = np.arange(0, b.n_pairs, 173)
    ro, ao = T.oracle_parallel(b.subset(sub), threads=min(16, os.cpu_count() or 1))
    T.compare(r2[sub].copy(), a2, ro, ao, what="streamed (ASCII, sub-ranges, masks) vs oracle")


def test_gpu_mid_width_bands_eight_lanes_per_alignment():
    """banded_sw with half-widths of 25-51 (indels of 24-50 bases, and narrower bands that double into that range): the warp band
    kernel's eight-lanes-per-alignment schedule (swb_bandwarp.cuh, Q = 8), with band doubling in place, under indelPost's penalty grid"""
    from gpuutil import gpu_align

    for seed, kw in ((71, dict(read_len=(150, 250), win_len=(500, 700), max_indel=50, grid=True)),
                     (72, dict(read_len=(120, 250), win_len=(400, 512), max_indel=40, grid=True, low_complexity=0.2, sub_rate=0.03)),
                     (73, dict(read_len=250, win_len=1000, max_indel=30, go=3, ge=1))):
        b = T.make_pairs(5000, seed=seed, **kw)
        rg, ag, tm = gpu_align(b)
        ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
        T.compare(rg, ag, ro, ao, what=f"mid-width bands seed {seed} vs oracle")


def test_gpu_sandwich_certificate_adversarial():
    """the sandwich sweep (swb_fast.cuh, SW = 1) certifies 8-bit-final results whose scores pass 128+go+ge; on a set built so that
    the 8-bit pass really deviates from Gotoh (insertions opened around score 128) every result must still equal the oracle's --
    a deviating pair that got certified would show up here -- and most of the zone must stay off the exact path"""
    from gpuutil import gpu_align

    b = T.make_sandwich_adversarial_pairs(40000, seed=5)
    rg, ag, tm = gpu_align(b)
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    T.compare(rg, ag, ro, ao, what="sandwich adversarial set vs oracle")
    # how many pairs really deviate: 8-bit result (score_size 2) vs plain Gotoh (score_size 1, 16-bit kernel only)
    import dataclasses
    r16, _ = T.oracle_parallel(dataclasses.replace(b, score_size=1), threads=min(16, os.cpu_count() or 1))
    dev = np.zeros(b.n_pairs, dtype=bool)
    for f in ("score1", "ref_end1", "read_end1", "ref_begin1", "read_begin1"):
        dev |= ro[f] != r16[f]
    zone = tm["n_sw_certified"] + tm["n_sw_rejected"]
    assert zone > 0.5 * b.n_pairs, (zone, b.n_pairs)
    assert dev.sum() > 0, "the generator no longer produces deviating pairs"
    assert tm["n_sw_certified"] <= zone - 0.5 * dev.sum(), (tm, int(dev.sum()))      # deviating pairs are rejected (forward or reverse)
    assert tm["n_sw_certified"] > 0.7 * zone, tm


def test_gpu_sandwich_short_read_fuzz():
    """30-100 bp reads x 300 bp windows under indelPost's penalty grid, N bases, junk tails: the zone the sandwich serves, bit-exact"""
    from gpuutil import gpu_align

    for seed, kw in ((31, dict(read_len=(30, 100), win_len=300, grid=True, max_indel=10)),
                     (32, dict(read_len=(44, 84), win_len=(120, 400), grid=True, max_indel=6, n_rate=0.01, junk_tail=0.2)),
                     (33, dict(read_len=(60, 84), win_len=300, grid=True, max_indel=3, low_complexity=0.3))):
        b = T.make_pairs(20000, seed=seed, **kw)
        rg, ag, tm = gpu_align(b)
        ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
        T.compare(rg, ag, ro, ao, what=f"short-read fuzz seed {seed} vs oracle")
        assert tm["n_sw_certified"] + tm["n_sw_rejected"] > 0


def test_gpu_overflow_verification_by_sandwich():
    """pairs whose CIGAR certificate fails (insertion opened around score 128, provisional 16-bit result) are settled by the sandwich
    lower bound; results stay bit-exact and most verifications never reach the exact 8-bit pass"""
    from gpuutil import gpu_align

    b = T.make_overflow_zone_pairs(30000, seed=9)
    rg, ag, tm = gpu_align(b)
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    T.compare(rg, ag, ro, ao, what="overflow zone vs oracle")
    assert tm["n_sw_verified"] > 0, tm


def _packed(b, bits):
    """the same batch with its sequence tables packed for SWB_SEQ_PACKED4 / PACKED2 (include/swb200.h)"""
    import copy
    from indelpost_b200.batch import pack_table

    q = copy.copy(b)
    q.reads, q.read_off = pack_table(b.reads, b.read_off, b.read_len, bits=bits)
    q.windows, q.win_off = pack_table(b.windows, b.win_off, b.win_len, bits=bits)
    q.reads = q.reads.view(np.int8); q.windows = q.windows.view(np.int8)
    q.seq_encoding = 2 if bits == 4 else 3
    return q


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [4, 2])
def test_gpu_packed_input_equals_codes(bits, monkeypatch):
    """4-bit and 2-bit packed sequence tables (unpacked on the device) give the results of the one-code-per-byte tables: single
    pass (ragged lengths, sub-ranges, N bases for 4 bits) and streamed path (pieces unpacked behind their copies)"""
    from gpuutil import gpu_align

    b = T.make_pairs(3000, (31, 150), (180, 400), seed=61, grid=True, n_rate=0.01 if bits == 4 else 0.0)
    rng = np.random.default_rng(8)
    wl = b.win_len[b.pair_win]
    b.ref_beg = rng.integers(0, 20, size=b.n_pairs).astype(np.int32)
    b.ref_len = (wl - b.ref_beg - rng.integers(0, 20, size=b.n_pairs)).astype(np.int32)
    r0, a0, _ = gpu_align(b)
    r1, a1, tm1 = gpu_align(_packed(b, bits))
    T.compare(r1, a1, r0, a0, what=f"{bits}-bit packed vs codes (single pass)")
    ro, ao = T.oracle().align_batch(b)
    T.compare(r1, a1, ro, ao, what=f"{bits}-bit packed vs oracle")
    big = T.make_pairs_fast(300000, 101, 251, seed=12, reads_per_window=20)
    big = big.subset(np.random.default_rng(4).permutation(big.n_pairs))      # upload frontier jumps around
    r2, a2, tm2 = gpu_align(_packed(big, bits))                              # >= 262144 pairs: streamed
    monkeypatch.setenv("SWB200_NO_PIPELINE", "1")
    r3, a3, tm3 = gpu_align(big)
    T.compare(r2, a2, r3, a3, what=f"{bits}-bit packed streamed vs codes single pass")
    assert tm2["h2d_bytes"] < tm3["h2d_bytes"] * (0.75 if bits == 4 else 0.55)


def test_multi_gpu_aligner_shards_and_stitches():
    """host-side sharding over several contexts (here: two contexts on the available device(s))"""
    import torch

    from indelpost_b200.sharding import MultiGpuAligner

    ndev = max(1, torch.cuda.device_count())
    devs = [0, 1 % ndev, 0][: 3 if ndev == 1 else 2]
    # windows up to 1000 columns and bands up to the widest class: every kernel family needs its opt-in shared-memory size on
    # EVERY device of the process (function attributes are per device)
    b = T.make_pairs(3000, (40, 150), (100, 1000), seed=81, reads_per_window=30, grid=True, n_rate=0.003, max_indel=25)
    m = MultiGpuAligner(devs)
    try:
        r, a = m.align(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext,
                       mat=b.mat, n=b.n, score_size=2, flag=1)
    finally:
        m.close()
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    T.compare(r.view(T.RESULT_DTYPE), a, ro, ao, what="MultiGpuAligner vs oracle")


def _random_matrix_batch(n, seed, n_pairs=600, rl=(30, 140), wl=(80, 300), vmax=6):
    rng = np.random.default_rng(seed)
    mat = rng.integers(-vmax, 1, size=(n, n)).astype(np.int8)
    mat = np.minimum(mat, mat.T)
    for i in range(n):
        mat[i, i] = rng.integers(1, vmax + 1)
    wins, reads = [], []
    for p in range(n_pairs):
        W = rng.integers(0, n, size=int(rng.integers(wl[0], wl[1] + 1)), dtype=np.int8)
        L = int(min(rng.integers(rl[0], rl[1] + 1), W.shape[0] - 5))
        s = int(rng.integers(0, W.shape[0] - L))
        r = W[s:s + L].copy()
        m = rng.random(L) < 0.05
        r[m] = rng.integers(0, n, size=int(m.sum()), dtype=np.int8)
        if rng.random() < 0.4:
            x = int(rng.integers(3, L - 3)); k = int(rng.integers(1, 6))
            r = np.concatenate([r[:x], r[x + k:]]) if rng.random() < 0.5 else np.concatenate([r[:x], rng.integers(0, n, size=k, dtype=np.int8), r[x:]])
        wins.append(W); reads.append(r.astype(np.int8))
    idx = np.arange(n_pairs, dtype=np.int32)
    b = T.batch_from_lists(reads, wins, idx, idx, rng.integers(2, 8, size=n_pairs), rng.integers(0, 3, size=n_pairs))
    b.mat = mat.reshape(-1)
    b.n = n
    return b


@pytest.mark.parametrize("n,vmax", [(4, 5), (5, 7), (8, 6), (21, 4), (5, 20)], ids=["n4", "n5", "n8", "n21-protein-like", "n5-big-scores"])
def test_gpu_general_substitution_matrices(n, vmax):
    """ssw_init takes any n x n matrix (ssw.h:86): fast path for |mat| <= 7 and windows of codes 0..3, exact path otherwise"""
    from gpuutil import gpu_align

    b = _random_matrix_batch(n, seed=400 + n + vmax, vmax=vmax)
    ro, ao = T.oracle().align_batch(b)
    rg, ag, _ = gpu_align(b)
    T.compare(rg, ag, ro, ao, what=f"matrix n={n} vmax={vmax}")


def test_gpu_long_reads_and_windows_beyond_fast_path_limits():
    from gpuutil import gpu_align

    cfgs = [dict(n_pairs=120, read_len=(257, 700), win_len=(800, 1500), seed=501, max_indel=30),      # reads longer than the fast path's 256 rows
            dict(n_pairs=40, read_len=(100, 250), win_len=(12000, 14000), seed=502, max_indel=10),     # windows beyond its shared-memory column budget
            dict(n_pairs=60, read_len=(342, 400), win_len=(500, 700), seed=503, max_indel=5)]          # max(mat)*L > 1023
    for cfg in cfgs:
        b = T.make_pairs(**cfg)
        ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
        rg, ag, tm = gpu_align(b)
        T.compare(rg, ag, ro, ao, what=f"long {cfg}")


@pytest.mark.parametrize("flag,filters,filterd", [(0, 0, 0), (8, 0, 0), (2, 420, 0), (4, 0, 148), (6, 400, 150), (1, 0, 0), (9, 0, 0), (3, 430, 0)])
def test_gpu_flag_and_filter_variants(flag, filters, filterd):
    """ssw_align's flag bits (ssw.c:816-824, 872, 894): score only / begin positions only / CIGAR only above a score
    filter or below a length filter.  Pairs that skip the traceback still have to pass the 8-bit/16-bit escalation check."""
    from gpuutil import gpu_align

    b = T.make_pairs(2500, (60, 150), 400, seed=600 + flag, max_indel=12, junk_tail=0.1)
    b.flag, b.filters, b.filterd = flag, filters, filterd
    ro, ao = T.oracle_parallel(b, threads=min(16, os.cpu_count() or 1))
    rg, ag, _ = gpu_align(b)
    T.compare(rg, ag, ro, ao, what=f"flag={flag} filters={filters} filterd={filterd}")
    if flag in (2, 4, 6, 3):
        assert 0 < int((ro["cigar_len"] > 0).sum()) < b.n_pairs      # the filters really split the batch


def test_gpu_mixed_window_lengths_long_windows_on_fast_path():
    """windows longer than 1024 columns keep the fast path (column bests in global memory); one batch mixes them
    with short windows and several read-length buckets"""
    from gpuutil import gpu_align

    parts = [T.make_pairs(300, 150, 400, seed=701), T.make_pairs(120, 250, (1500, 2600), seed=702, max_indel=60),
             T.make_pairs(200, (40, 100), 300, seed=703), T.make_pairs(60, 150, (5000, 9000), seed=704)]
    reads, wins, go, ge = [], [], [], []
    for b in parts:
        for p in range(b.n_pairs):
            reads.append(b.reads[b.read_off[p]:b.read_off[p] + b.read_len[p]])
            w = b.pair_win[p]
            wins.append(b.windows[b.win_off[w]:b.win_off[w] + b.win_len[w]])
            go.append(int(b.gap_open[p])); ge.append(int(b.gap_ext[p]))
    idx = np.arange(len(reads), dtype=np.int32)
    perm = np.random.default_rng(7).permutation(len(reads))
    bb = T.batch_from_lists(reads, wins, idx[perm], idx[perm], np.array(go)[perm], np.array(ge)[perm])
    bb.mat = T.dna_matrix(3, 2)
    ro, ao = T.oracle_parallel(bb, threads=min(16, os.cpu_count() or 1))
    rg, ag, tm = gpu_align(bb)
    T.compare(rg, ag, ro, ao, what="mixed window lengths")
    assert tm["n_fast"] > 0.8 * bb.n_pairs


# ---------------------------------------------------------------------------------------------
# CIGAR -> indel records on the device (include/swb200.h section 3, swb_indels.cuh)
# ---------------------------------------------------------------------------------------------

def _oracle_records(cigar, rs, qs):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import indels_oracle as O

    recs, rend = O.indel_records(cigar, rs, qs)
    return [((n << 4) | (1 if op == "I" else 2), ri, qi, po) for op, n, ri, qi, po in recs], rend


def test_gpu_indels_reproduce_reference_golden():
    """findall_indels through the device kernel equals the dicts the unmodified reference produced"""
    from collections import namedtuple

    from indelpost_b200 import localn

    Aln = namedtuple("Aln", "CIGAR reference_start read_start")
    with open(os.path.join(GOLDEN_DIR, "indels.json")) as fh:
        cases = json.load(fh)["cases"]
    for snv in (False, True):
        sel = [c for c in cases if c["report_snvs"] == snv]
        outs = localn.findall_indels_many([Aln(c["cigar"], c["reference_start"], c["read_start"]) for c in sel], [c["genome_aln_pos"] for c in sel],
                                          [c["ref_seq"] for c in sel], [c["read_seq"] for c in sel], report_snvs=snv)
        for c, out in zip(sel, outs):
            indels, snvs = out if snv else (out, None)
            assert indels == c["indels"], c["cigar"]
            assert snvs == c["snvs"], c["cigar"]
    c = cases[0]
    one = localn.findall_indels(Aln(c["cigar"], c["reference_start"], c["read_start"]), c["genome_aln_pos"], c["ref_seq"], c["read_seq"], report_snvs=c["report_snvs"])
    assert (one[0] if c["report_snvs"] else one) == c["indels"]


def test_gpu_indels_fuzz_against_oracle():
    """random op strings (adjacent gap runs in every order, gaps at both ends, empty CIGARs) through swb_indels_from_cigars"""
    from gpuutil import aligner
    from indelpost_b200 import localn

    rng = np.random.default_rng(77)
    cigars, rs, qs = [], [], []
    for k in range(20000):
        n = int(rng.integers(0, 9))
        toks = []
        for _ in range(n):
            op = "MID"[int(rng.choice(3, p=[0.4, 0.3, 0.3]))]
            if toks and toks[-1][-1] == op == "M":
                continue
            toks.append(f"{int(rng.integers(1, 40))}{op}")
        cigars.append("".join(toks))
        rs.append(int(rng.integers(0, 50)))
        qs.append(int(rng.integers(0, 20)))
    packed = [localn._pack_cigar(c) for c in cigars]
    clen = np.array([p.shape[0] for p in packed], dtype=np.int32)
    coff = np.zeros(len(cigars), dtype=np.int64)
    coff[1:] = np.cumsum(clen[:-1])
    off, cnt, rend, recs = aligner().indels_from_cigars(np.concatenate(packed), coff, clen, rs, qs)
    assert int(cnt.sum()) == recs.shape[0]
    for k, c in enumerate(cigars):
        want, want_end = _oracle_records(c, rs[k], qs[k])
        got = [(int(r["cigar_op"]), int(r["ref_idx"]), int(r["read_idx"]), int(r["pos_off"])) for r in recs[int(off[k]): int(off[k]) + int(cnt[k])]]
        assert got == want, (c, got, want)
        assert int(rend[k]) == want_end, c
        assert all(int(r["pair"]) == k for r in recs[int(off[k]): int(off[k]) + int(cnt[k])])


def test_gpu_indels_of_resident_alignments():
    """swb_indels on the alignments a compute just produced equals the oracle walk over the downloaded CIGARs"""
    from gpuutil import aligner

    b = T.make_pairs(5000, (60, 150), 400, seed=91, grid=True, max_indel=12)
    a = aligner()
    n = a.upload(b.reads, b.read_off, b.read_len, b.windows, b.win_off, b.win_len, b.pair_read, b.pair_win, b.gap_open, b.gap_ext, mat=b.mat, n=5, score_size=2, flag=1)
    a.compute()
    res, arena = a.download(n)
    off, cnt, rend, recs = a.indels(n)
    n_events = 0
    for p in range(n):
        cg = T.cigar_string(arena, int(res["cigar_off"][p]), int(res["cigar_len"][p])) if res["cigar_len"][p] > 0 else ""
        want, want_end = _oracle_records(cg, int(res["ref_begin1"][p]), int(res["read_begin1"][p]))
        got = [(int(r["cigar_op"]), int(r["ref_idx"]), int(r["read_idx"]), int(r["pos_off"])) for r in recs[int(off[p]): int(off[p]) + int(cnt[p])]]
        assert got == want, (p, cg)
        assert int(rend[p]) == want_end
        n_events += len(want)
    assert n_events > 1000
