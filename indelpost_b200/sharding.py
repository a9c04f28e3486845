"""Host-side sharding of a pair batch over several GPUs.

Pairs are independent and results return per pair, so multi-GPU needs no collective: the batch is cut into
contiguous ranges with balanced DP work (sum of readLen * windowLen), each range goes to its own swb_ctx
(one per device, one host thread each), and the per-shard results are stitched back in pair order with the CIGAR
offsets rebased.  Contiguous ranges keep a locus' reads together, so its window is uploaded to one device only.

`shard_bounds` / `stitch` are pure numpy and are what the torchrun path of bench.py and the gloo tests use;
`MultiGpuAligner` is the in-process (threads) variant for one host driving several GPUs.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Sequence, Tuple

import ctypes as C

import numpy as np

from . import _lib as L


def pair_cells(read_len, win_len, pair_read, pair_win, ref_len=None, ref_beg=None) -> np.ndarray:
    """nominal DP cells of every pair: readLen * searched window length (the tables are gathered as they are: no conversion pass
    over a table that may be far longer than the pair list)"""
    read_len, win_len = np.asarray(read_len), np.asarray(win_len)
    rl = read_len[np.asarray(pair_read).clip(0, max(len(read_len) - 1, 0))].astype(np.int64)
    if ref_len is not None:
        wl = np.asarray(ref_len).astype(np.int64)
    else:
        wl = win_len[np.asarray(pair_win).clip(0, max(len(win_len) - 1, 0))].astype(np.int64)
        if ref_beg is not None:
            wl = wl - np.asarray(ref_beg).astype(np.int64)
    return rl * np.maximum(wl, 0)


def shard_bounds(cells: np.ndarray, n_shards: int) -> List[Tuple[int, int]]:
    """contiguous [p0, p1) ranges whose cell sums are as equal as a contiguous cut allows"""
    n = int(cells.shape[0])
    n_shards = max(1, int(n_shards))
    if n == 0:
        return [(0, 0)] * n_shards
    csum = np.cumsum(cells.astype(np.float64))
    total = csum[-1]
    cuts = [0]
    for k in range(1, n_shards):
        target = total * k / n_shards
        cuts.append(int(np.searchsorted(csum, target, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.array(cuts))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(n_shards)]


def stitch(parts: Sequence[Tuple[np.ndarray, np.ndarray]]) -> Tuple[np.ndarray, np.ndarray]:
    """concatenate per-shard (results, cigar_arena) in shard order, rebasing cigar_off"""
    res = [np.array(r, copy=True) for r, _ in parts]
    base = 0
    for r, (_, a) in zip(res, parts):
        if r.shape[0]:
            r["cigar_off"] = np.where(r["cigar_len"] > 0, r["cigar_off"] + base, r["cigar_off"])
        base += int(a.shape[0])
    arena = np.concatenate([a for _, a in parts]) if parts else np.zeros(0, np.uint32)
    return (np.concatenate(res) if res else np.zeros(0, L.RESULT_DTYPE)), arena


def slice_pairs(arrs: dict, p0: int, p1: int) -> dict:
    """per-pair arrays restricted to [p0, p1)"""
    out = dict(arrs)
    for k in ("pair_read", "pair_win", "gap_open", "gap_ext", "ref_beg", "ref_len", "mask_len"):
        if out.get(k) is not None:
            out[k] = np.ascontiguousarray(out[k][p0:p1])
    return out


def slice_table(blob, off, length, idx, seq_encoding: int = L.SWB_SEQ_CODES, scratch: dict = None, key: str = ""):
    """the part of a sequence table the (valid) indices in `idx` refer to: (blob slice, rebased offsets, lengths, first index).
    A shard uploads these instead of the whole table (each GPU gets only the sequences its pairs touch).  Entries are
    taken as an index RANGE [min, max], so out-of-range pair indices still fail inside the library like they do unsharded."""
    n = int(np.asarray(length).shape[0])
    idx = np.asarray(idx)
    if idx.shape[0] == 0 or n == 0:
        return blob[:0], np.zeros(0, np.int64), np.zeros(0, np.int32), 0
    i0, i1 = int(idx.min()), int(idx.max()) + 1          # two reductions; the masked path only when an index is out of range
    if i0 < 0 or i1 > n:
        ok = idx[(idx >= 0) & (idx < n)]
        if ok.shape[0] == 0:
            return blob[:0], np.zeros(0, np.int64), np.zeros(0, np.int32), 0
        i0, i1 = int(ok.min()), int(ok.max()) + 1
    o = np.ascontiguousarray(np.asarray(off)[i0:i1], dtype=np.int64)
    ln = np.ascontiguousarray(np.asarray(length)[i0:i1], dtype=np.int32)
    shift = {L.SWB_SEQ_PACKED4: 1, L.SWB_SEQ_PACKED2: 2}.get(int(seq_encoding), 0)
    if scratch is not None:                                  # reuse the rebased-offset array across calls (no first-touch page faults)
        buf = scratch.get(key)
        if buf is None or buf.shape[0] < i1 - i0:
            buf = scratch[key] = np.empty(max(1024, (i1 - i0) + (i1 - i0) // 8), dtype=np.int64)
        out_off = buf[: i1 - i0]
    else:
        out_off = np.empty(i1 - i0, dtype=np.int64)
    ext = (C.c_int64 * 2)()
    # one pass in C (GIL released): byte extent [b0, b1) of the entries and their offsets rebased to b0; -1: a negative offset / length
    if L.load().swb_slice_table(o.ctypes.data, ln.ctypes.data, i1 - i0, shift, out_off.ctypes.data, ext) != 0:
        return blob, o, ln, i0                                              # let the library report the bad table
    return blob[int(ext[0]):int(ext[1])], out_off, ln, i0


def sampled_bounds(read_len, win_len, pair_read, pair_win, ref_len, ref_beg, n_shards: int, stride: int = 16) -> List[Tuple[int, int]]:
    """shard_bounds over every `stride`-th pair (a 16th of the gathers and of the prefix sum): the cut positions are good to
    `stride` pairs, which is far below what a shard's balance needs"""
    n = int(np.asarray(pair_read).shape[0])
    if n_shards <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(1, n_shards) - 1)
    if n < 64 * stride:
        stride = 1
    sl = slice(0, n, stride)
    cells = pair_cells(read_len, win_len, pair_read[sl], pair_win[sl], None if ref_len is None else ref_len[sl], None if ref_beg is None else ref_beg[sl])
    b = shard_bounds(cells, n_shards)
    cuts = [min(n, p0 * stride) for p0, _ in b] + [n]
    cuts[0] = 0
    return [(cuts[k], cuts[k + 1]) for k in range(n_shards)]


class MultiGpuAligner:
    """one BatchAligner per device, one host thread per device, no collective.

    Every shard (a contiguous range of pairs, balanced by nominal DP cells) uploads only the table slices its pairs refer to
    and writes its records straight into its slice of ONE pinned result array; CIGARs go to the shard's own region of one
    pinned arena (`cigar_off` of the records is rebased to the arena start, so `arena[off : off + len]` works as for a single
    GPU -- the regions are not packed, unused entries between them are never referenced).  The returned arrays are views of
    buffers owned by this object, valid until its next `align`."""

    def __init__(self, devices: Sequence[int]):
        from .batch import BatchAligner

        self.devices = list(devices)
        self.aligners = [BatchAligner(d) for d in self.devices]
        self.pool = ThreadPoolExecutor(max_workers=len(self.devices))
        self._res = None
        self._arena = None
        self._scratch = [dict() for _ in self.devices]

    def close(self):
        for a in self.aligners:
            a.close()
        self.pool.shutdown(wait=False)
        for b in (self._res, self._arena):
            if b is not None:
                b.close()
        self._res = self._arena = None

    def _out(self, n_pairs: int, arena_entries: int):
        need_r = max(1, n_pairs) * L.RESULT_DTYPE.itemsize
        if self._res is None or self._res.nbytes < need_r:
            if self._res is not None:
                self._res.close()
            self._res = L.PinnedBuffer(need_r + need_r // 8)
        need_a = max(64, arena_entries) * 4
        if self._arena is None or self._arena.nbytes < need_a:
            if self._arena is not None:
                self._arena.close()
            self._arena = L.PinnedBuffer(need_a + need_a // 8)
        return self._res.view(L.RESULT_DTYPE, n_pairs), self._arena.view(np.uint32, self._arena.nbytes // 4)

    def align(self, reads, read_off, read_len, windows, win_off, win_len, pair_read, pair_win, gap_open, gap_ext,
              ref_beg=None, ref_len=None, mask_len=None, copy: bool = True, cigar_per_pair: int = 8, **kw):
        arrs = dict(pair_read=np.asarray(pair_read, np.int32), pair_win=np.asarray(pair_win, np.int32),
                    gap_open=np.asarray(gap_open, np.uint8), gap_ext=np.asarray(gap_ext, np.uint8),
                    ref_beg=None if ref_beg is None else np.asarray(ref_beg, np.int32),
                    ref_len=None if ref_len is None else np.asarray(ref_len, np.int32),
                    mask_len=None if mask_len is None else np.asarray(mask_len, np.int32))
        n = int(arrs["pair_read"].shape[0])
        bounds = sampled_bounds(read_len, win_len, arrs["pair_read"], arrs["pair_win"], arrs["ref_len"], arrs["ref_beg"], len(self.aligners))
        enc = int(kw.get("seq_encoding", L.SWB_SEQ_CODES))
        reads_a, windows_a = np.asarray(reads), np.asarray(windows)
        # arena regions: a fixed budget per pair (grown and retried if a shard reports it needs more)
        per = max(2, int(cigar_per_pair))
        while True:
            abase = [0]
            for p0, p1 in bounds:
                abase.append(abase[-1] + max(64, per * (p1 - p0)))
            res, arena = self._out(n, abase[-1])

            def run(k):
                p0, p1 = bounds[k]
                if p1 <= p0:
                    return 0
                s = slice_pairs_view(arrs, p0, p1)
                # only the table slices this shard refers to travel to its GPU; pair indices are rebased by subtraction (an index
                # outside the caller's tables stays outside the slice)
                rb, ro, rl, r0 = slice_table(reads_a, read_off, read_len, s["pair_read"], enc, self._scratch[k], "r")
                wb, wo, wl, w0 = slice_table(windows_a, win_off, win_len, s["pair_win"], enc, self._scratch[k], "w")
                pr = s["pair_read"] - np.int32(r0) if r0 else s["pair_read"]
                pw = s["pair_win"] - np.int32(w0) if w0 else s["pair_win"]
                used = self.aligners[k].align_into(res[p0:p1], arena[abase[k]:abase[k + 1]], rb, ro, rl, wb, wo, wl, pr, pw,
                                                   s["gap_open"], s["gap_ext"], ref_beg=s["ref_beg"], ref_len=s["ref_len"], mask_len=s["mask_len"], **kw)
                if used >= 0 and abase[k]:
                    self.aligners[k].lib.swb_rebase_cigar_offsets(res[p0:p1].ctypes.data, p1 - p0, abase[k])      # in C, GIL released
                return used

            used = list(self.pool.map(run, range(len(self.aligners))))
            if min(used) >= 0:
                break
            per = max(per * 2, max((-u + (p1 - p0) - 1) // max(1, p1 - p0) for u, (p0, p1) in zip(used, bounds) if u < 0) + 1)
        a = arena[: abase[-1]]
        return (res.copy(), a.copy()) if copy else (res, a)


def slice_pairs_view(arrs: dict, p0: int, p1: int) -> dict:
    """per-pair arrays restricted to [p0, p1) as views (a contiguous range of a contiguous array is contiguous)"""
    return {k: (None if v is None else v[p0:p1]) for k, v in arrs.items()}
