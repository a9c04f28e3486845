"""Shared helpers for the parity tests: synthetic pair generators, ctypes bindings of the
CPU oracle (oracle/_build/libssw_oracle.so), of the compiled reference (oracle/_ref/libssw_ref.so,
when present) and batch containers with the same SoA layout as include/swb200.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module's oracle bindings; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libssw_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libssw_ref.so")

# same layout as swb_result (include/swb200.h) / orc_batch_result / ref_batch_result
RESULT_DTYPE = np.dtype(
    [
        ("score1", "<u2"),
        ("score2", "<u2"),
        ("ref_begin1", "<i4"),
        ("ref_end1", "<i4"),
        ("read_begin1", "<i4"),
        ("read_end1", "<i4"),
        ("ref_end2", "<i4"),
        ("cigar_len", "<i4"),
        ("flag", "<u2"),
        ("status", "<u2"),
        ("cigar_off", "<i8"),
    ],
    align=True,
)
assert RESULT_DTYPE.itemsize == 40, RESULT_DTYPE.itemsize

RESULT_FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "flag", "status")

_LUT = np.full(256, 4, dtype=np.int8)
for _ch, _v in (("A", 0), ("a", 0), ("C", 1), ("c", 1), ("G", 2), ("g", 2), ("T", 3), ("t", 3), ("U", 0), ("u", 0)):
    _LUT[ord(_ch)] = _v


def encode_dna(s) -> np.ndarray:
    """sswpy.pyx:16-29 DNA_BASE_LUT."""
    if isinstance(s, str):
        s = s.encode()
    return _LUT[np.frombuffer(s, dtype=np.uint8)]


def dna_matrix(match: int = 3, mismatch: int = 2) -> np.ndarray:
    """sswpy.pyx:306-336 buildDNAScoreMatrix (5x5, zeros for N)."""
    m = np.zeros((5, 5), dtype=np.int8)
    for i in range(4):
        for j in range(4):
            m[i, j] = match if i == j else -mismatch
    return m.reshape(-1)


@dataclass
class Batch:
    """Host-side SoA batch (mirror of swb_batch)."""

    reads: np.ndarray  # int8 blob
    read_off: np.ndarray  # int64 [n_reads]
    read_len: np.ndarray  # int32 [n_reads]
    windows: np.ndarray  # int8 blob
    win_off: np.ndarray  # int64 [n_windows]
    win_len: np.ndarray  # int32 [n_windows]
    pair_read: np.ndarray  # int32 [n_pairs]
    pair_win: np.ndarray  # int32 [n_pairs]
    gap_open: np.ndarray  # uint8 [n_pairs]
    gap_ext: np.ndarray  # uint8 [n_pairs]
    ref_beg: np.ndarray | None = None
    ref_len: np.ndarray | None = None
    mask_len: np.ndarray | None = None
    mat: np.ndarray = field(default_factory=dna_matrix)
    n: int = 5
    score_size: int = 2
    flag: int = 1
    filters: int = 0
    filterd: int = 0
    seq_encoding: int = 0

    @property
    def n_pairs(self) -> int:
        return int(self.pair_read.shape[0])

    @property
    def n_reads(self) -> int:
        return int(self.read_len.shape[0])

    @property
    def n_windows(self) -> int:
        return int(self.win_len.shape[0])

    def cells(self) -> int:
        rl = self.read_len[self.pair_read].astype(np.int64)
        if self.ref_len is not None:
            wl = self.ref_len.astype(np.int64)
        else:
            wl = self.win_len[self.pair_win].astype(np.int64)
            if self.ref_beg is not None:
                wl = wl - self.ref_beg
        return int((rl * wl).sum())

    def subset(self, idx) -> "Batch":
        idx = np.asarray(idx)
        sel = lambda a: None if a is None else np.ascontiguousarray(a[idx])
        return Batch(
            self.reads, self.read_off, self.read_len, self.windows, self.win_off, self.win_len,
            sel(self.pair_read), sel(self.pair_win), sel(self.gap_open), sel(self.gap_ext),
            sel(self.ref_beg), sel(self.ref_len), sel(self.mask_len),
            self.mat, self.n, self.score_size, self.flag, self.filters, self.filterd, self.seq_encoding,
        )


def batch_from_lists(reads, windows, pair_read, pair_win, go, ge, **kw) -> Batch:
    """reads / windows: lists of int8 code arrays (or ASCII bytes when seq_encoding=1)."""

    def blob(seqs):
        arrs = [np.frombuffer(s, dtype=np.int8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.int8) for s in seqs]
        lens = np.array([a.shape[0] for a in arrs], dtype=np.int32)
        off = np.zeros(len(arrs), dtype=np.int64)
        if len(arrs) > 1:
            off[1:] = np.cumsum(lens[:-1], dtype=np.int64)
        data = np.concatenate(arrs) if arrs and sum(lens) else np.zeros(0, dtype=np.int8)
        return np.ascontiguousarray(data), off, lens

    r, ro, rl = blob(reads)
    w, wo, wl = blob(windows)
    n = len(pair_read)
    go = np.full(n, go, dtype=np.uint8) if np.isscalar(go) else np.asarray(go, dtype=np.uint8)
    ge = np.full(n, ge, dtype=np.uint8) if np.isscalar(ge) else np.asarray(ge, dtype=np.uint8)
    for k in ("ref_beg", "ref_len", "mask_len"):
        if kw.get(k) is not None:
            kw[k] = np.asarray(kw[k], dtype=np.int32)
    return Batch(r, ro, rl, w, wo, wl, np.asarray(pair_read, dtype=np.int32), np.asarray(pair_win, dtype=np.int32), go, ge, **kw)


# --------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------

def make_pairs(
    n_pairs: int,
    read_len=150,
    win_len=400,
    *,
    seed: int = 1,
    reads_per_window: int = 1,
    max_indel: int = 10,
    sub_rate: float = 0.01,
    n_rate: float = 0.0,
    win_n_rate: float = 0.0,
    go=3,
    ge=1,
    low_complexity: float = 0.0,
    junk_tail: float = 0.0,
    match: int = 3,
    mismatch: int = 2,
    grid: bool = False,
) -> Batch:
    """Config-2 style generator: per pair a read sampled from its window with one planted event
    (1/3 none, 1/3 deletion 1..max_indel, 1/3 insertion 1..max_indel), substitutions and optional N.

    read_len / win_len may be ints or (lo, hi) ranges.  reads_per_window > 1 shares windows
    (the realistic locus layout).  grid=True draws (go, ge) from indelPost's grid
    (varaln.pyx:1127-1143) plus go=len(read) (localn.pyx:255) and go=ge=len(read) (varaln.pyx:1228-1234).
    """
    rng = np.random.default_rng(seed)
    n_windows = (n_pairs + reads_per_window - 1) // reads_per_window

    def draw(spec, size):
        if isinstance(spec, (tuple, list)):
            return rng.integers(spec[0], spec[1] + 1, size=size).astype(np.int32)
        return np.full(size, spec, dtype=np.int32)

    wlen = draw(win_len, n_windows)
    woff = np.zeros(n_windows, dtype=np.int64)
    woff[1:] = np.cumsum(wlen[:-1], dtype=np.int64)
    windows = rng.integers(0, 4, size=int(wlen.sum()), dtype=np.int8)
    if low_complexity > 0:
        for w in np.nonzero(rng.random(n_windows) < low_complexity)[0]:
            unit = rng.integers(0, 4, size=int(rng.integers(1, 5)), dtype=np.int8)
            reps = int(wlen[w]) // len(unit) + 1
            windows[woff[w] : woff[w] + wlen[w]] = np.tile(unit, reps)[: wlen[w]]
    if win_n_rate > 0:
        windows[rng.random(windows.shape[0]) < win_n_rate] = 4

    rlen_target = draw(read_len, n_pairs)
    pair_win = (np.arange(n_pairs) // reads_per_window).astype(np.int32)
    reads = []
    rl_out = np.zeros(n_pairs, dtype=np.int32)
    kinds = rng.integers(0, 3, size=n_pairs)
    for p in range(n_pairs):
        w = pair_win[p]
        W = windows[woff[w] : woff[w] + wlen[w]]
        L = int(min(rlen_target[p], max(1, wlen[w] - max_indel - 1)))
        k = int(kinds[p])
        ev = int(rng.integers(1, max_indel + 1)) if max_indel > 0 else 0
        if k == 0 or ev == 0 or L < 4:
            span = L
            start = int(rng.integers(0, wlen[w] - span + 1))
            r = W[start : start + span].copy()
        elif k == 1:  # deletion: read skips ev reference bases
            span = L + ev
            start = int(rng.integers(0, max(1, wlen[w] - span + 1)))
            cut = int(rng.integers(1, L))
            r = np.concatenate([W[start : start + cut], W[start + cut + ev : start + span]])
        else:  # insertion: read has ev extra bases
            ev = min(ev, L - 2)
            span = L - ev
            start = int(rng.integers(0, wlen[w] - span + 1))
            cut = int(rng.integers(1, span))
            ins = rng.integers(0, 4, size=ev, dtype=np.int8)
            r = np.concatenate([W[start : start + cut], ins, W[start + cut : start + span]])
        r = r.astype(np.int8)
        if sub_rate > 0 and r.shape[0]:
            m = rng.random(r.shape[0]) < sub_rate
            r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.int8)) % 4
        if junk_tail > 0 and rng.random() < junk_tail and r.shape[0] > 8:
            t = int(rng.integers(3, max(4, r.shape[0] // 3)))
            if rng.random() < 0.5:
                r[-t:] = rng.integers(0, 4, size=t, dtype=np.int8)
            else:
                r[:t] = rng.integers(0, 4, size=t, dtype=np.int8)
        if n_rate > 0 and r.shape[0]:
            r[rng.random(r.shape[0]) < n_rate] = 4
        reads.append(r)
        rl_out[p] = r.shape[0]
    roff = np.zeros(n_pairs, dtype=np.int64)
    roff[1:] = np.cumsum(rl_out[:-1], dtype=np.int64)
    blob = np.concatenate(reads) if reads else np.zeros(0, dtype=np.int8)

    if grid:
        table = np.array([(3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0), (0, 0), (0, 0)], dtype=np.int64)
        pick = rng.integers(0, 8, size=n_pairs)
        gos = table[pick, 0].copy()
        ges = table[pick, 1].copy()
        m6 = pick == 6  # mut aligner: go = len(read), ge = 1
        gos[m6] = rl_out[m6]
        ges[m6] = 1
        m7 = pick == 7  # is_perfect_match: go = ge = len(read)
        gos[m7] = rl_out[m7]
        ges[m7] = rl_out[m7]
        gos = (gos % 256).astype(np.uint8)
        ges = (ges % 256).astype(np.uint8)
    else:
        gos = np.full(n_pairs, go, dtype=np.uint8) if np.isscalar(go) else np.asarray(go, dtype=np.uint8)
        ges = np.full(n_pairs, ge, dtype=np.uint8) if np.isscalar(ge) else np.asarray(ge, dtype=np.uint8)

    return Batch(
        np.ascontiguousarray(blob), roff, rl_out, np.ascontiguousarray(windows), woff, wlen,
        np.arange(n_pairs, dtype=np.int32), pair_win, gos, ges, mat=dna_matrix(match, mismatch),
    )


# --------------------------------------------------------------------------------------
# ctypes bindings of the checkers
# --------------------------------------------------------------------------------------

def _ptr(a, ct):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ct))


def build_oracle(force: bool = False) -> None:
    """Compile oracle/ (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "ssw_oracle.c")):
        subprocess.run(["make", "-C", ORACLE_DIR, "_build/libssw_oracle.so"], check=True, capture_output=True)
    subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=False, capture_output=True)


_BATCH_ARGS = [
    C.POINTER(C.c_int8), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
    C.POINTER(C.c_int8), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
    C.POINTER(C.c_int32), C.POINTER(C.c_int32),
    C.POINTER(C.c_int32), C.POINTER(C.c_int32),
    C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
    C.POINTER(C.c_int8), C.c_int32, C.c_int8,
    C.c_uint8, C.c_uint16, C.c_int32,
    C.c_void_p, C.POINTER(C.c_uint32), C.c_int64,
]


def _batch_call_args(b: Batch):
    return [
        _ptr(b.reads, C.c_int8), _ptr(b.read_off, C.c_int64), _ptr(b.read_len, C.c_int32),
        _ptr(b.windows, C.c_int8), _ptr(b.win_off, C.c_int64), _ptr(b.win_len, C.c_int32),
        _ptr(b.pair_read, C.c_int32), _ptr(b.pair_win, C.c_int32),
        _ptr(b.ref_beg, C.c_int32), _ptr(b.ref_len, C.c_int32),
        _ptr(b.gap_open, C.c_uint8), _ptr(b.gap_ext, C.c_uint8), _ptr(b.mask_len, C.c_int32),
        _ptr(np.ascontiguousarray(b.mat, dtype=np.int8), C.c_int8), b.n, b.score_size,
        b.flag, b.filters, b.filterd,
    ]


def _decode_ascii(b: Batch) -> Batch:
    if b.seq_encoding == 0:
        return b
    import dataclasses

    return dataclasses.replace(b, reads=_LUT[b.reads.view(np.uint8)], windows=_LUT[b.windows.view(np.uint8)], seq_encoding=0)


class _CpuChecker:
    def __init__(self, path: str, symbol: str, has_range: bool):
        self.lib = C.CDLL(path)
        self.fn = getattr(self.lib, symbol)
        self.fn.restype = C.c_int64
        self.fn.argtypes = ([C.c_int32, C.c_int32] if has_range else [C.c_int32]) + _BATCH_ARGS
        self.has_range = has_range

    def align_batch(self, b: Batch, first: int = 0, count: int | None = None):
        b = _decode_ascii(b)
        n = b.n_pairs
        count = n - first if count is None else count
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = int(b.read_len[b.pair_read[first : first + count]].astype(np.int64).sum() + 2 * count + 16)
        arena = np.zeros(cap, dtype=np.uint32)
        if self.has_range:
            head = [first, count]
        else:
            assert first == 0 and count == n
            head = [n]
        used = self.fn(*head, *_batch_call_args(b), res.ctypes.data_as(C.c_void_p), _ptr(arena, C.c_uint32), cap)
        assert used <= cap
        return res, arena[:used]


_oracle = None
_ref = None


def oracle() -> _CpuChecker:
    global _oracle
    if _oracle is None:
        build_oracle()
        _oracle = _CpuChecker(ORACLE_SO, "orc_align_batch", False)
    return _oracle


def have_ref() -> bool:
    if not os.path.exists(REF_SO):
        build_oracle()
    return os.path.exists(REF_SO)


def reference() -> _CpuChecker:
    """The reference's own ssw.c (oracle/_ref).  Callers must check have_ref() first."""
    global _ref
    if _ref is None:
        _ref = _CpuChecker(REF_SO, "ref_align_batch", True)
    return _ref


# --------------------------------------------------------------------------------------
# comparison
# --------------------------------------------------------------------------------------

_OPS = "MIDNSHP=X"


def cigar_string(arena: np.ndarray, off: int, ln: int):
    if ln == 0:
        return None
    return "".join(f"{int(v) >> 4}{_OPS[int(v) & 15] if (int(v) & 15) <= 8 else 'M'}" for v in arena[off : off + ln])


def compare(res_a, arena_a, res_b, arena_b, what="", limit=5):
    """Assert bit-exact equality of every s_align field and every CIGAR op."""
    assert res_a.shape == res_b.shape
    bad = np.zeros(res_a.shape[0], dtype=bool)
    for f in RESULT_FIELDS:
        bad |= res_a[f] != res_b[f]
    ok = ~bad
    if ok.any():
        # CIGAR content
        la = res_a["cigar_len"].astype(np.int64)
        idx = np.nonzero(ok & (la > 0))[0]
        if idx.size:
            oa = res_a["cigar_off"][idx]
            ob = res_b["cigar_off"][idx]
            ln = la[idx]
            tot = int(ln.sum())
            rep = np.repeat(np.arange(idx.size), ln)
            within = np.arange(tot) - np.repeat(np.cumsum(ln) - ln, ln)
            va = arena_a[oa[rep] + within]
            vb = arena_b[ob[rep] + within]
            neq = va != vb
            if neq.any():
                bad[idx[np.unique(rep[neq])]] = True
    nbad = int(bad.sum())
    if nbad:
        lines = []
        for p in np.nonzero(bad)[0][:limit]:
            a = {f: int(res_a[f][p]) for f in RESULT_FIELDS}
            b = {f: int(res_b[f][p]) for f in RESULT_FIELDS}
            a["cigar"] = cigar_string(arena_a, int(res_a["cigar_off"][p]), int(res_a["cigar_len"][p]))
            b["cigar"] = cigar_string(arena_b, int(res_b["cigar_off"][p]), int(res_b["cigar_len"][p]))
            lines.append(f"pair {p}:\n   A={a}\n   B={b}")
        raise AssertionError(f"{what}: {nbad}/{res_a.shape[0]} pairs differ\n" + "\n".join(lines))
    return True


def make_pairs_fast(n_pairs: int, read_len: int = 150, win_len: int = 400, *, seed: int = 1, reads_per_window: int = 1,
                    max_indel: int = 10, sub_rate: float = 0.01, go: int = 3, ge: int = 1, match: int = 3, mismatch: int = 2,
                    chunk: int = 65536) -> Batch:
    """Vectorised config-2 generator for bench-sized batches (1 M pairs in a few seconds): fixed read
    and window lengths, per pair one planted event (1/3 none, 1/3 deletion, 1/3 insertion of
    1..max_indel bp) plus substitutions — the same distribution as make_pairs()."""
    rng = np.random.default_rng(seed)
    n_windows = (n_pairs + reads_per_window - 1) // reads_per_window
    windows = rng.integers(0, 4, size=(n_windows, win_len), dtype=np.int8)
    reads = np.empty((n_pairs, read_len), dtype=np.int8)
    pair_win = (np.arange(n_pairs) // reads_per_window).astype(np.int32)
    L = read_len
    k = np.arange(L, dtype=np.int32)[None, :]
    for c0 in range(0, n_pairs, chunk):
        c1 = min(n_pairs, c0 + chunk)
        m = c1 - c0
        kind = rng.integers(0, 3, size=m)
        ev = rng.integers(1, max_indel + 1, size=m).astype(np.int32) if max_indel > 0 else np.zeros(m, np.int32)
        ev = np.where(kind == 0, 0, ev)
        ev = np.where(kind == 2, np.minimum(ev, L - 2), ev)
        span = np.where(kind == 1, L + ev, np.where(kind == 2, L - ev, L)).astype(np.int32)
        start = (rng.random(m) * (win_len - span + 1)).astype(np.int32)
        cut = (1 + rng.random(m) * (np.where(kind == 2, span, L) - 1)).astype(np.int32)
        cut = np.maximum(1, np.minimum(cut, np.where(kind == 2, span, L) - 1))
        is_del = (kind == 1)[:, None]
        is_ins = (kind == 2)[:, None]
        evc = ev[:, None]
        cutc = cut[:, None]
        # source index in the window for every read position
        src = start[:, None] + k + np.where(is_del & (k >= cutc), evc, 0) - np.where(is_ins & (k >= cutc + evc), evc, 0)
        src = np.clip(src, 0, win_len - 1)
        r = windows[pair_win[c0:c1, None], src]
        inserted = is_ins & (k >= cutc) & (k < cutc + evc)
        rnd = rng.integers(0, 4, size=(m, L), dtype=np.int8)
        r = np.where(inserted, rnd, r)
        if sub_rate > 0:
            sub = rng.random((m, L)) < sub_rate
            r = np.where(sub, (r + 1 + (rnd % 3)) % 4, r)
        reads[c0:c1] = r.astype(np.int8)
    roff = (np.arange(n_pairs, dtype=np.int64) * L)
    woff = (np.arange(n_windows, dtype=np.int64) * win_len)
    return Batch(
        np.ascontiguousarray(reads.reshape(-1)), roff, np.full(n_pairs, L, dtype=np.int32),
        np.ascontiguousarray(windows.reshape(-1)), woff, np.full(n_windows, win_len, dtype=np.int32),
        np.arange(n_pairs, dtype=np.int32), pair_win,
        np.full(n_pairs, go, dtype=np.uint8), np.full(n_pairs, ge, dtype=np.uint8), mat=dna_matrix(match, mismatch),
    )


def make_window_edge_pairs(n_pairs: int, seed: int = 1) -> Batch:
    """Reads anchored at the very start (or end) of their window with an indel close to the anchored end, mixed
    gap penalties with gap_extension > 0: the reverse rectangle (ssw.c:875-886) then has fewer columns than rows
    and the optimal start lies off the main diagonal -- the geometry the banded reverse pass (swb_revband.cuh)
    must get right (band columns beyond the rectangle, first rows of the band, competing equal-score starts)."""
    rng = np.random.default_rng(seed)
    grid = [(3, 1), (5, 1), (4, 1), (6, 2), (2, 1), (7, 3)]
    wins, reads, go, ge = [], [], [], []
    for p in range(n_pairs):
        wl = int(rng.integers(120, 420))
        W = rng.integers(0, 4, size=wl, dtype=np.int8)
        L = int(rng.integers(40, min(160, wl - 30)))
        ev = int(rng.integers(1, 14))
        kind = int(rng.integers(0, 3))
        at_end = rng.random() < 0.3
        if kind == 0:      # insertion near the anchored end
            span = L - ev
            start = int(rng.integers(0, 5)) if not at_end else wl - span - int(rng.integers(0, 5))
            cut = int(rng.integers(1, max(2, min(24, span - 1)))) if not at_end else span - int(rng.integers(1, max(2, min(24, span - 1))))
            ins = rng.integers(0, 4, size=ev, dtype=np.int8) if rng.random() < 0.6 else W[max(0, start + cut - ev) : start + cut][:ev].copy()
            r = np.concatenate([W[start : start + cut], ins, W[start + cut : start + span]])
        elif kind == 1:    # deletion near the anchored end
            span = L + ev
            start = int(rng.integers(0, 5)) if not at_end else wl - span - int(rng.integers(0, 5))
            cut = int(rng.integers(1, 24)) if not at_end else L - int(rng.integers(1, 24))
            r = np.concatenate([W[start : start + cut], W[start + cut + ev : start + span]])
        else:              # both, a few bases apart
            ev2 = int(rng.integers(1, 8))
            span = L - ev + ev2
            start = int(rng.integers(0, 5)) if not at_end else wl - span - int(rng.integers(0, 5))
            cut = int(rng.integers(2, 16))
            cut2 = cut + int(rng.integers(3, 12))
            r = np.concatenate([W[start : start + cut], rng.integers(0, 4, size=ev, dtype=np.int8), W[start + cut : start + cut2], W[start + cut2 + ev2 : start + span]])
        r = r.astype(np.int8)
        m = rng.random(r.shape[0]) < 0.015
        r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.int8)) % 4
        g = grid[int(rng.integers(0, len(grid)))]
        wins.append(W); reads.append(r); go.append(g[0]); ge.append(g[1])
    return batch_from_lists(reads, wins, np.arange(n_pairs), np.arange(n_pairs), np.array(go), np.array(ge))


def make_overflow_zone_pairs(n_pairs: int, seed: int = 1) -> Batch:
    """Targeted stress for the 8-bit/16-bit escalation decision (ssw.c:842-847): reads whose best
    alignment needs an insertion placed where the running score is around 128 (the zone in which the
    signed lazy-F exit test of ssw.c:311 can drop a vertical gap), so that whether the 8-bit pass
    overflows depends on that quirk.  Mixed gap penalties from indelPost's grid."""
    rng = np.random.default_rng(seed)
    grid = [(3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)]
    wins, reads, go, ge = [], [], [], []
    for p in range(n_pairs):
        wl = int(rng.integers(250, 420))
        W = rng.integers(0, 4, size=wl, dtype=np.int8)
        L = int(rng.integers(86, 170))
        q = int(rng.integers(30, min(100, L - 10)))           # insertion point: left flank score ~ 3q
        k = int(rng.integers(1, 14))
        span = L - k
        start = int(rng.integers(0, wl - span))
        ins = rng.integers(0, 4, size=k, dtype=np.int8)
        r = np.concatenate([W[start:start + q], ins, W[start + q:start + span]]).astype(np.int8)
        nsub = int(rng.integers(0, 4))
        for _ in range(nsub):
            x = int(rng.integers(0, r.shape[0]))
            r[x] = (r[x] + 1 + rng.integers(0, 3)) % 4
        if rng.random() < 0.2:                                 # second event
            x = int(rng.integers(5, r.shape[0] - 5))
            r = np.concatenate([r[:x], r[x + int(rng.integers(1, 4)):]]) if rng.random() < 0.5 else np.concatenate([r[:x], rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.int8), r[x:]])
        g = grid[int(rng.integers(0, 6))]
        wins.append(W); reads.append(r.astype(np.int8)); go.append(g[0]); ge.append(g[1])
    idx = np.arange(n_pairs, dtype=np.int32)
    b = batch_from_lists(reads, wins, idx, idx, go, ge)
    b.mat = dna_matrix(3, 2)
    return b


def make_sandwich_adversarial_pairs(n_pairs: int, seed: int = 1) -> Batch:
    """Adversarial set for the sandwich certificate of 8-bit-final results (swb_fast.cuh, SW = 1): short reads (62-83 bp, the 8-bit
    pass cannot overflow) with an insertion placed where the running score is around 128 + go, so that the signed lazy-F exit test
    (ssw.c:309-311) really drops vertical gaps and the 8-bit result differs from Gotoh for some pairs -- those must not be certified."""
    rng = np.random.default_rng(seed)
    grid = [(3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0)]
    wins, reads, go, ge = [], [], [], []
    for p in range(n_pairs):
        wl = int(rng.integers(150, 300))
        W = rng.integers(0, 4, size=wl, dtype=np.int8)
        L = int(rng.integers(62, 84))
        q = int(rng.integers(40, 52))                          # left flank scores ~120-156: the gap opens around 128 + go
        k = int(rng.integers(1, 10))
        span = L - k
        start = int(rng.integers(0, wl - span))
        ins = rng.integers(0, 4, size=k, dtype=np.int8)
        r = np.concatenate([W[start:start + q], ins, W[start + q:start + span]]).astype(np.int8)
        for _ in range(int(rng.integers(0, 3))):
            x = int(rng.integers(0, r.shape[0]))
            r[x] = (r[x] + 1 + rng.integers(0, 3)) % 4
        g = grid[int(rng.integers(0, 6))]
        wins.append(W); reads.append(r); go.append(g[0]); ge.append(g[1])
    idx = np.arange(n_pairs, dtype=np.int32)
    b = batch_from_lists(reads, wins, idx, idx, go, ge)
    b.mat = dna_matrix(3, 2)
    return b


def oracle_parallel(b: Batch, threads: int = 8):
    """oracle over a batch using several threads (ctypes releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor

    parts = [p for p in np.array_split(np.arange(b.n_pairs), threads) if p.shape[0]]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        outs = list(ex.map(lambda p: oracle().align_batch(b.subset(p)), parts))
    res = np.concatenate([o[0] for o in outs])
    shift = 0
    arenas = []
    pos = 0
    for (r, a), p in zip(outs, parts):
        res["cigar_off"][pos:pos + p.shape[0]] += shift
        shift += a.shape[0]
        pos += p.shape[0]
        arenas.append(a)
    return res, np.concatenate(arenas) if arenas else np.zeros(0, np.uint32)
