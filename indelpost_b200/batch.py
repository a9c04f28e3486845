"""Batched aligner: the new entry point that the sswpy-compatible layer adds.

`BatchAligner` owns one swb_ctx (= one GPU).  `align()` takes the SoA arrays of include/swb200.h's
swb_batch as numpy arrays and returns (results, cigar_arena) with results a structured array of
RESULT_DTYPE (every s_align field of reference ssw.h:55-66)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def dna_score_matrix(match_score: int = 2, mismatch_penalty: int = 2) -> np.ndarray:
    """sswpy.pyx:306-336 buildDNAScoreMatrix: 5x5, +match on the ACGT diagonal, -mismatch elsewhere,
    0 for anything involving N.  The arguments are narrowed to uint8 then int8 like the reference does."""
    m = np.zeros(25, dtype=np.int8)
    ms = np.array([match_score & 0xFF], dtype=np.uint8).view(np.int8)[0]
    mm = np.array([(-(mismatch_penalty & 0xFF)) & 0xFF], dtype=np.uint8).view(np.int8)[0]
    for i in range(4):
        for j in range(4):
            m[i * 5 + j] = ms if i == j else mm
    return m


def _c(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def pack_table(blob, off, length, bits: int = 4, ascii: bool = False, out=None, out_off=None):
    """Pack a sequence table (codes, or ASCII with ``ascii=True``) for ``seq_encoding=SWB_SEQ_PACKED4`` (bits=4) or
    ``SWB_SEQ_PACKED2`` (bits=2; no N): returns (packed_blob uint8, packed_off int64).  ``out`` / ``out_off`` may be
    preallocated (pinned) arrays.  include/swb200.h: swb_pack_table."""
    lib = L.load()
    blob = np.ascontiguousarray(blob).view(np.int8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    n = int(length.shape[0])
    per = 8 // bits
    need = int(((length.astype(np.int64) + per - 1) // per).sum())
    dst = out if out is not None else np.empty(max(need, 1), dtype=np.uint8)
    doff = out_off if out_off is not None else np.empty(max(n, 1), dtype=np.int64)
    if dst.nbytes < need or doff.shape[0] < n:
        raise ValueError("pack_table: output arrays too small")
    used = lib.swb_pack_table(blob.ctypes.data, off.ctypes.data, length.ctypes.data, n, 1 if ascii else 0, int(bits), dst.ctypes.data, doff.ctypes.data)
    if used < 0:
        raise ValueError("pack_table: a code does not fit the packing (2 bits hold A/C/G/T only) or an argument is bad")
    return dst[:used], doff[:n]


class _OutLease:
    """a (results, cigar arena) pair of pinned buffers on loan from a BatchAligner's pool"""

    def __init__(self, owner, n, cap):
        self.owner = owner
        need_r, need_a = max(1, n) * L.RESULT_DTYPE.itemsize, max(64, cap) * 4
        pool = owner._out_pool
        pick = None
        for i, (r, a) in enumerate(pool):
            if r.nbytes >= need_r and a.nbytes >= need_a:
                pick = i
                break
        if pick is not None:
            self.res, self.arena = pool.pop(pick)
        else:
            if len(pool) >= 4:                      # keep the pool bounded: drop the smallest idle pair
                j = min(range(len(pool)), key=lambda q: pool[q][0].nbytes + pool[q][1].nbytes)
                for b in pool.pop(j):
                    b.close()
            self.res = L.PinnedBuffer(need_r + need_r // 8)
            self.arena = L.PinnedBuffer(need_a + need_a // 8)

    def views(self, n):
        return self.res.view(L.RESULT_DTYPE, n), self.arena.view(np.uint32, self.arena.nbytes // 4)

    def grow_arena(self, cap):
        self.arena.close()
        self.arena = L.PinnedBuffer(cap * 4 + cap // 2)

    def release(self):
        if self.res is not None:
            if getattr(self.owner, "ctx", None):
                self.owner._out_pool.append((self.res, self.arena))
            else:
                self.res.close(); self.arena.close()
            self.res = self.arena = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class BatchAligner:
    def __init__(self, device: int = 0):
        self.lib = L.load()
        self.ctx = self.lib.swb_create(int(device))
        if not self.ctx:
            raise L.SwbError(self.lib.swb_last_error(None).decode())
        self.device = device
        self._keep = None
        self._res_pin = None      # pinned output staging, grown on demand and reused across calls
        self._arena_pin = None
        self._out_pool = []       # pinned (results, arena) buffer pairs handed out by align_leased() and returned by their lease

    def _out_buffers(self, n: int, cap: int):
        """pinned (cudaMallocHost) result / CIGAR buffers: D2H into pageable numpy memory would be staged and
        synchronous, which serialises the pipelined swb_align_batch"""
        need_r = max(1, n) * L.RESULT_DTYPE.itemsize
        if self._res_pin is None or self._res_pin.nbytes < need_r:
            if self._res_pin is not None:
                self._res_pin.close()
            self._res_pin = L.PinnedBuffer(need_r + need_r // 8)
        need_a = max(64, cap) * 4
        if self._arena_pin is None or self._arena_pin.nbytes < need_a:
            if self._arena_pin is not None:
                self._arena_pin.close()
            self._arena_pin = L.PinnedBuffer(need_a + need_a // 8)
        return self._res_pin.view(L.RESULT_DTYPE, n), self._arena_pin.view(np.uint32, self._arena_pin.nbytes // 4)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.swb_destroy(self.ctx)
            self.ctx = None
        for b in ("_res_pin", "_arena_pin"):
            if getattr(self, b, None) is not None:
                getattr(self, b).close()
                setattr(self, b, None)
        for r, a in getattr(self, "_out_pool", []):
            r.close(); a.close()
        self._out_pool = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _make_batch(self, reads, read_off, read_len, windows, win_off, win_len, pair_read, pair_win, gap_open, gap_ext,
                    ref_beg=None, ref_len=None, mask_len=None, mat=None, n=5, score_size=2, flag=1, filters=0, filterd=0,
                    seq_encoding=L.SWB_SEQ_CODES):
        arrs = dict(
            reads=_c(reads, np.int8), read_off=_c(read_off, np.int64), read_len=_c(read_len, np.int32),
            windows=_c(windows, np.int8), win_off=_c(win_off, np.int64), win_len=_c(win_len, np.int32),
            pair_read=_c(pair_read, np.int32), pair_win=_c(pair_win, np.int32),
            ref_beg=_c(ref_beg, np.int32), ref_len=_c(ref_len, np.int32),
            gap_open=_c(gap_open, np.uint8), gap_ext=_c(gap_ext, np.uint8), mask_len=_c(mask_len, np.int32),
            mat=_c(mat if mat is not None else dna_score_matrix(), np.int8),
        )
        b = L.SwbBatch()
        b.n_pairs = int(arrs["pair_read"].shape[0])
        b.n_reads = int(arrs["read_len"].shape[0])
        b.n_windows = int(arrs["win_len"].shape[0])
        # the C ABI trusts these sizes: a short array would be read past its end
        for k in ("pair_win", "gap_open", "gap_ext", "ref_beg", "ref_len", "mask_len"):
            if arrs[k] is not None and arrs[k].shape != (b.n_pairs,):
                raise ValueError(f"{k} must have one entry per pair ({b.n_pairs}), got shape {arrs[k].shape}")
        if arrs["read_off"].shape != (b.n_reads,) or arrs["win_off"].shape != (b.n_windows,):
            raise ValueError("read_off / read_len and win_off / win_len must have equal sizes")
        if arrs["mat"].size != int(n) * int(n):
            raise ValueError(f"mat must hold n*n = {int(n) * int(n)} entries")
        b.seq_encoding = int(seq_encoding)
        for k, a in arrs.items():
            setattr(b, k, None if a is None else a.ctypes.data)
        b.n = int(n)
        b.score_size = int(score_size)
        b.flag = int(flag)
        b.filters = int(filters)
        b.filterd = int(filterd)
        return b, arrs

    def _err(self):
        return self.lib.swb_last_error(self.ctx).decode()

    def align(self, *args, cigar_cap: int | None = None, copy: bool = True, **kw):
        """one-shot: host arrays in, (results, cigar_arena) out (H2D + kernels + D2H).

        Outputs land in pinned buffers owned by the aligner; with copy=False the returned arrays are views
        of those buffers, valid until the next call on this aligner."""
        b, keep = self._make_batch(*args, **kw)
        n = b.n_pairs
        cap = int(cigar_cap if cigar_cap is not None else max(64, 8 * n))
        used = C.c_int64(0)
        while True:
            res, arena = self._out_buffers(n, cap)
            cap = arena.shape[0]
            rc = self.lib.swb_align_batch(self.ctx, C.byref(b), res.ctypes.data, arena.ctypes.data, cap, C.byref(used))
            if rc == -2:
                cap = int(used.value) + 64
                continue
            if rc != 0:
                raise L.SwbError(self._err())
            a = arena[: used.value]
            return (res.copy(), a.copy()) if copy else (res, a)

    def align_into(self, res: np.ndarray, arena: np.ndarray, *args, **kw) -> int:
        """one-shot into caller-owned (ideally pinned) output arrays: `res` (RESULT_DTYPE, one record per pair) and `arena`
        (uint32).  Returns the number of arena entries used, or -(needed) if the arena is too small (nothing valid written).
        This is what a sharder uses to let every GPU write its slice of ONE result array (no stitching copies)."""
        b, keep = self._make_batch(*args, **kw)
        if res.shape[0] < b.n_pairs or res.dtype != L.RESULT_DTYPE or not res.flags.c_contiguous or arena.dtype != np.uint32:
            raise ValueError("align_into: res must be a contiguous RESULT_DTYPE array with one record per pair, arena uint32")
        used = C.c_int64(0)
        rc = self.lib.swb_align_batch(self.ctx, C.byref(b), res.ctypes.data, arena.ctypes.data, int(arena.shape[0]), C.byref(used))
        if rc == -2:
            return -int(used.value)
        if rc != 0:
            raise L.SwbError(self._err())
        return int(used.value)

    def align_leased(self, *args, cigar_cap: int | None = None, **kw):
        """like align(), but the outputs land in pinned buffers taken from a pool and stay valid for as long as the returned
        lease object lives (the buffers go back to the pool when it is released) -- no copy of the records"""
        b, keep = self._make_batch(*args, **kw)
        n = b.n_pairs
        cap = int(cigar_cap if cigar_cap is not None else max(64, 8 * n))
        used = C.c_int64(0)
        lease = _OutLease(self, n, cap)
        while True:
            res, arena = lease.views(n)
            rc = self.lib.swb_align_batch(self.ctx, C.byref(b), res.ctypes.data, arena.ctypes.data, arena.shape[0], C.byref(used))
            if rc == -2:
                lease.grow_arena(int(used.value) + 64)
                continue
            if rc != 0:
                lease.release()
                raise L.SwbError(self._err())
            return res, arena[: used.value], lease

    # split form (device-resident timing)
    def upload(self, *args, **kw):
        b, keep = self._make_batch(*args, **kw)
        self._keep = (b, keep)
        if self.lib.swb_upload(self.ctx, C.byref(b)) != 0:
            raise L.SwbError(self._err())
        return b.n_pairs

    def compute(self):
        if self.lib.swb_compute(self.ctx) != 0:
            raise L.SwbError(self._err())

    def download(self, n_pairs: int, cigar_cap: int | None = None):
        res = np.zeros(n_pairs, dtype=L.RESULT_DTYPE)
        cap = int(cigar_cap if cigar_cap is not None else max(64, 16 * n_pairs))
        used = C.c_int64(0)
        while True:
            arena = np.zeros(cap, dtype=np.uint32)
            rc = self.lib.swb_download(self.ctx, res.ctypes.data, arena.ctypes.data, cap, C.byref(used))
            if rc == -2:
                cap = int(used.value) + 64
                continue
            if rc != 0:
                raise L.SwbError(self._err())
            return res, arena[: used.value]

    # ------------------------------------------------------------------
    # CIGAR -> indel records (include/swb200.h section 3; reference localn.pyx:542-621 + utilities.pyx:360-401)
    def _indel_call(self, n, fn):
        off = np.zeros(max(n, 1), dtype=np.int64)
        cnt = np.zeros(max(n, 1), dtype=np.int32)
        rend = np.zeros(max(n, 1), dtype=np.int32)
        cap = max(64, 2 * n)
        used = C.c_int64(0)
        while True:
            recs = np.zeros(cap, dtype=L.INDEL_DTYPE)
            rc = fn(off, cnt, rend, recs, cap, used)
            if rc == -2:
                cap = int(used.value) + 16
                continue
            if rc != 0:
                raise L.SwbError(self._err())
            return off[:n], cnt[:n], rend[:n], recs[: used.value]

    def indels(self, n_pairs: int):
        """indel records of the alignments this aligner holds on the device (after compute() or a streamed / single-pass
        align()): (indel_off, indel_cnt, read_end, records) -- the records of pair p are records[off[p] : off[p] + cnt[p]]"""
        return self._indel_call(n_pairs, lambda off, cnt, rend, recs, cap, used: self.lib.swb_indels(
            self.ctx, off.ctypes.data, cnt.ctypes.data, rend.ctypes.data, recs.ctypes.data, cap, C.byref(used)))

    def indels_from_cigars(self, cigar_arena, cigar_off, cigar_len, ref_start, read_start):
        """the same for caller-supplied alignments (BAM-packed CIGARs in an arena)"""
        arena = np.ascontiguousarray(cigar_arena, dtype=np.uint32)
        coff = np.ascontiguousarray(cigar_off, dtype=np.int64)
        clen = np.ascontiguousarray(cigar_len, dtype=np.int32)
        rs = np.ascontiguousarray(ref_start, dtype=np.int32)
        qs = np.ascontiguousarray(read_start, dtype=np.int32)
        n = int(coff.shape[0])
        return self._indel_call(n, lambda off, cnt, rend, recs, cap, used: self.lib.swb_indels_from_cigars(
            self.ctx, n, arena.ctypes.data, int(arena.shape[0]), coff.ctypes.data, clen.ctypes.data, rs.ctypes.data, qs.ctypes.data,
            off.ctypes.data, cnt.ctypes.data, rend.ctypes.data, recs.ctypes.data, cap, C.byref(used)))

    def timing(self) -> dict:
        t = L.SwbTiming()
        self.lib.swb_get_timing(self.ctx, C.byref(t))
        return t.as_dict()
