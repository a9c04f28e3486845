"""CPU tests of the Python host layer's native helpers (indelpost_b200/csrc/swbhost.c) and of the prefetch block index --
no GPU, no compute: staged inputs and lazily built Alignment tuples must be byte-for-byte what the pure-Python
formulation gives."""
import numpy as np
import pytest

from indelpost_b200 import _lib as L
from indelpost_b200 import _swbhost as H
from indelpost_b200 import sswpy


def test_gather_copies_str_and_bytes_into_the_staging_blob():
    seqs = ["ACGT", b"TTGCA", "", "nnA", "ACGU" * 50]
    total = H.total_len(seqs)
    assert total == sum(len(s) for s in seqs)
    blob = np.zeros(total + 8, np.int8)
    off = np.zeros(len(seqs), np.int64)
    ln = np.zeros(len(seqs), np.int32)
    end = H.gather(seqs, blob.ctypes.data, total, off.ctypes.data, ln.ctypes.data, 0)
    assert end == total
    raw = blob.view(np.uint8).tobytes()
    for s, o, l in zip(seqs, off, ln):
        want = s if isinstance(s, bytes) else s.encode()
        assert raw[o: o + l] == want
    # too small a buffer: nothing is copied, the needed capacity comes back as -(need) - 1
    assert H.gather(seqs, blob.ctypes.data, total - 1, off.ctypes.data, ln.ctypes.data, 0) == -total - 1
    # a table large enough for the threaded copy (GIL released, shares split by byte count)
    rng = np.random.default_rng(3)
    big = [bytes(rng.integers(65, 90, size=int(k), dtype=np.uint8)) for k in rng.integers(0, 700, size=40000)]
    tb = sum(map(len, big))
    bb = np.zeros(tb, np.int8); bo = np.zeros(len(big), np.int64); bl = np.zeros(len(big), np.int32)
    assert H.gather(big, bb.ctypes.data, tb, bo.ctypes.data, bl.ctypes.data, 0) == tb
    assert bb.view(np.uint8).tobytes() == b"".join(big)
    assert bo.tolist() == np.concatenate([[0], np.cumsum([len(x) for x in big])[:-1]]).tolist()
    with pytest.raises(TypeError):
        H.gather(["ACGT", 5], blob.ctypes.data, total, off.ctypes.data, ln.ctypes.data, 0)
    # non-ASCII text is encoded as UTF-8 like obj_to_cstr_len does (sswpy.pyx:45-55)
    assert H.total_len(["Aé"]) == 3


def _records(rng, n):
    res = np.zeros(n, L.RESULT_DTYPE)
    lens = rng.integers(0, 6, n)
    off = np.concatenate([[0], np.cumsum(lens)[:-1]])
    arena = ((rng.integers(1, 3000, int(lens.sum())).astype(np.uint32)) << 4) | rng.integers(0, 12, int(lens.sum())).astype(np.uint32)
    for f in ("score1", "score2"):
        res[f] = rng.integers(0, 60000, n)
    for f in ("ref_begin1", "ref_end1", "read_begin1", "read_end1"):
        res[f] = rng.integers(-1, 100000, n)
    res["cigar_len"], res["cigar_off"] = lens, off
    return res, arena


def test_lazy_alignment_list_matches_the_python_formatting():
    rng = np.random.default_rng(3)
    res, arena = _records(rng, 500)
    al = sswpy.AlignmentList(res, arena)
    want = []
    for r in res:
        ops = arena[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_len"])]
        cig = sswpy.cigar_to_string(ops) if len(ops) else None
        want.append(sswpy.Alignment(cig, int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"])))
    assert len(al) == 500
    assert al.tolist() == want and list(al) == want and al == want
    assert al[7] == want[7] and al[-1] == want[-1] and al[10:13] == want[10:13]
    assert isinstance(al[0], sswpy.Alignment) and al[0].optimal_score == want[0].optimal_score
    with pytest.raises(IndexError):
        al[500]


def test_prefetch_blocks_index_pairs_grid_and_len_penalties():
    sswpy.clear_prefetched()
    rng = np.random.default_rng(5)
    reads = [b"A" * 100, b"C" * 150, b"G" * 300]
    wins = [b"T" * 300, b"TA" * 200]
    grid = ((3, 1), (4, 0), ("len", 1), ("len", "len"))
    n = len(reads) * len(wins)
    res, arena = _records(rng, n * len(grid))
    al = sswpy.AlignmentList(res, arena)
    rids, wids = [sswpy.seq_id(r) for r in reads], [sswpy.seq_id(w) for w in wins]
    mkey = b"m"
    sswpy.register_block(al, rids, wids, None, None, grid, mkey, cross=True)
    for r in range(3):
        for w in range(2):
            p = r * 2 + w
            L_ = len(reads[r])
            assert sswpy.prefetched(mkey, rids[r], wids[w], 3, 1, L_) == al[0 * n + p]
            assert sswpy.prefetched(mkey, rids[r], wids[w], 4, 0, L_) == al[1 * n + p]
            assert sswpy.prefetched(mkey, rids[r], wids[w], L_ & 0xFF, 1, L_) == al[2 * n + p]          # 300 narrows to 44 (uint8_t, sswpy.pyx:214-219)
            assert sswpy.prefetched(mkey, rids[r], wids[w], L_ & 0xFF, L_ & 0xFF, L_) == al[3 * n + p]
            assert sswpy.prefetched(mkey, rids[r], wids[w], 5, 1, L_) is None
            assert sswpy.prefetched(b"other", rids[r], wids[w], 3, 1, L_) is None
    hit = sswpy.prefetched(mkey, rids[0], wids[0], 3, 1, 100)
    assert sswpy.prefetched(mkey, rids[0], wids[0], 3, 1, 100) is hit          # built once, kept
    # explicit pair lists
    pr, pw = np.array([2, 0], np.int32), np.array([1, 1], np.int32)
    res2, arena2 = _records(rng, 2 * 2)
    al2 = sswpy.AlignmentList(res2, arena2)
    sswpy.register_block(al2, rids, wids, pr, pw, ((7, 2), (9, 9)), b"k", cross=False)
    assert sswpy.prefetched(b"k", rids[2], wids[1], 7, 2, 300) == al2[0]
    assert sswpy.prefetched(b"k", rids[0], wids[1], 9, 9, 100) == al2[3]
    assert sswpy.prefetched(b"k", rids[1], wids[1], 7, 2, 150) is None
    # ids are never reused: an aligner holding an old id cannot hit a block registered after clear_prefetched()
    old = rids[0]
    sswpy.clear_prefetched()
    assert sswpy.seq_id(b"brand new") > old
    assert sswpy.prefetched(mkey, rids[0], wids[0], 3, 1, 100) is None
    sswpy.clear_prefetched()
