"""sswpy-compatible host layer over libswb200 (mirrors reference indelpost/sswpy.pyx).

Same names, argument meaning and error behaviour as the reference's Cython module:
  * ``SSW(match_score=2, mismatch_penalty=2)`` with ``setRead`` / ``setReference`` / ``align``
    (sswpy.pyx:99-304) -- every ``align`` goes through the C ABI's ``ssw_init`` / ``ssw_align``
    (include/swb200.h), i.e. runs on the GPU; there is no CPU implementation here.
  * ``Alignment`` 7-field named tuple (sswpy.pyx:85-94).
  * ``force_align`` / ``format_force_align`` (sswpy.pyx:339-396).
  * NEW: ``align_batch`` -- many read x reference pairs in one call (swb_align_batch).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence, Union

import numpy as np

from . import _lib as L
from .batch import BatchAligner, dna_score_matrix

STR_T = Union[str, bytes]


class Alignment(NamedTuple):  # sswpy.pyx:85-94
    CIGAR: Optional[str]
    optimal_score: int
    sub_optimal_score: int
    reference_start: int
    reference_end: int
    read_start: int
    read_end: int


_LUT = np.full(256, 4, dtype=np.int8)  # sswpy.pyx:16-25 (entries >= 128 are outside the reference's table: treated as 4)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 0)):
    _LUT[ord(_ch)] = _v
    _LUT[ord(_ch.lower())] = _v

_OPS = "MIDNSHP=X"


def _to_bytes(o: STR_T) -> bytes:  # obj_to_cstr_len, sswpy.pyx:45-55
    if isinstance(o, bytes):
        return o
    if isinstance(o, str):
        return o.encode("utf8")
    raise TypeError("expected str or bytes")


def _c_int(v, name="value") -> int:
    v = int(v)
    if not (-(2**31) <= v < 2**31):
        raise OverflowError(f"{name} too large to convert to C int")
    return v


def cigar_to_string(ops: np.ndarray) -> str:
    """sswpy.pyx:283-289: "%d%s" per BAM-packed op (len<<4|op, MAPSTR "MIDNSHP=X", ssw.h:171-190)"""
    return "".join(f"{int(v) >> 4}{_OPS[int(v) & 15] if (int(v) & 15) <= 8 else 'M'}" for v in ops)


class SSW:
    """Drop-in for reference ``sswpy.SSW`` (sswpy.pyx:99-337)."""

    def __init__(self, match_score: int = 2, mismatch_penalty: int = 2):
        self._lib = L.load()
        self.score_matrix = dna_score_matrix(_c_int(match_score), _c_int(mismatch_penalty))  # buildDNAScoreMatrix
        self._mkey = self.score_matrix.tobytes()
        self.read = None
        self.reference = None
        self._read_arr = None
        self._ref_arr = None
        self._profile = None
        self.read_length = 0
        self.ref_length = 0
        # results of this aligner's recent calls: indelPost repeats alignments with identical inputs (update_read_info,
        # pileup.pyx:849, re-runs what retarget computed at pileup.pyx:647 -- SURVEY.md 8f item 1); a hit skips the GPU round trip
        self._memo = {}

    def __del__(self):  # __dealloc__, sswpy.pyx:135-147
        try:
            if self._profile:
                self._lib.init_destroy(self._profile)
                self._profile = None
        except Exception:
            pass

    def setRead(self, read: STR_T):  # sswpy.pyx:149-178
        raw = _to_bytes(read)
        arr = np.ascontiguousarray(_LUT[np.frombuffer(raw, dtype=np.uint8)])
        if self._profile:
            self._lib.init_destroy(self._profile)
            self._profile = None
        self.read = read
        self._read_arr = arr
        self.read_length = len(raw)
        self._profile = self._lib.ssw_init(arr.ctypes.data, len(raw), self.score_matrix.ctypes.data, 5, 2)

    def setReference(self, reference: STR_T):  # sswpy.pyx:180-197
        raw = _to_bytes(reference)
        self._ref_arr = np.ascontiguousarray(_LUT[np.frombuffer(raw, dtype=np.uint8)])
        self._ref_key = self._ref_arr.tobytes()
        self.reference = reference
        self.ref_length = len(raw)
        self._memo.clear()

    def align(self, gap_open: int = 3, gap_extension: int = 1, start_idx: int = 0, end_idx: int = 0) -> Alignment:
        """sswpy.pyx:227-304 (same checks, same order, same messages)"""
        gap_open = _c_int(gap_open, "gap_open")
        gap_extension = _c_int(gap_extension, "gap_extension")
        start_idx = int(start_idx)
        end_idx = int(end_idx)
        if start_idx < 0 or end_idx < 0:
            raise ValueError("negative indexing not supported")
        if end_idx > self.ref_length or start_idx > self.ref_length:
            raise ValueError(
                "start_idx: {} or end_idx: {} can't be greater than ref_length: {}".format(start_idx, end_idx, self.ref_length)
            )
        end_idx_final = self.ref_length if end_idx == 0 else end_idx
        search_length = end_idx_final - start_idx
        if self.reference is None:
            raise ValueError("call setReference first")
        if not self._profile:
            raise ValueError("Must set profile first")
        key = (self._read_arr.tobytes(), gap_open & 0xFF, gap_extension & 0xFF, start_idx, search_length)
        hit = self._memo.get(key)
        if hit is None and _PREFETCHED:
            hit = _PREFETCHED.get((self._mkey, self._ref_key) + key)          # filled by prefetch_alignments()
        if hit is not None:
            return hit
        mask_len = self.read_length // 2  # align_c, sswpy.pyx:209-211
        mask_len = 15 if mask_len < 15 else mask_len
        ref_ptr = self._ref_arr.ctypes.data + start_idx
        # gap penalties are C ints narrowed to uint8_t by ssw_align's prototype (sswpy.pyx:214-219)
        res = self._lib.ssw_align(self._profile, ref_ptr, search_length, gap_open & 0xFF, gap_extension & 0xFF, 1, 0, 0, mask_len)
        if not res:
            raise ValueError("Problem Running alignment, see stdout")
        r = res.contents
        cigar = None
        if r.cigar:
            ops = np.ctypeslib.as_array(r.cigar, shape=(r.cigarLen,)) if r.cigarLen > 0 else np.zeros(0, np.uint32)
            cigar = cigar_to_string(ops)
        out = Alignment(cigar, r.score1, r.score2, r.ref_begin1, r.ref_end1, r.read_begin1, r.read_end1)
        self._lib.align_destroy(res)
        if len(self._memo) >= 256:
            self._memo.clear()
        self._memo[key] = out
        return out


# ---------------------------------------------------------------------------------------------
# batched entry point
# ---------------------------------------------------------------------------------------------

_default_aligners: dict = {}


def _aligner(device: int) -> BatchAligner:
    a = _default_aligners.get(device)
    if a is None:
        a = _default_aligners[device] = BatchAligner(device)
    return a


def _blob(seqs: Sequence[STR_T]):
    raws = [_to_bytes(s) for s in seqs]
    lens = np.fromiter((len(r) for r in raws), dtype=np.int32, count=len(raws))
    off = np.zeros(len(raws), dtype=np.int64)
    if len(raws) > 1:
        np.cumsum(lens[:-1], out=off[1:])
    data = np.frombuffer(b"".join(raws), dtype=np.int8) if raws else np.zeros(0, np.int8)
    return data, off, lens


def align_batch(
    reads: Sequence[STR_T],
    references: Sequence[STR_T],
    pair_read: Sequence[int],
    pair_ref: Sequence[int],
    gap_open=3,
    gap_extension=1,
    start_idx=None,
    end_idx=None,
    match_score: int = 2,
    mismatch_penalty: int = 2,
    device: int = 0,
    aligner: Optional[BatchAligner] = None,
):
    """Align ``reads[pair_read[k]]`` against ``references[pair_ref[k]]`` for every k in ONE GPU batch.

    Each result equals what ``SSW(match_score, mismatch_penalty)`` + ``setReference`` + ``setRead`` +
    ``align(gap_open[k], gap_extension[k], start_idx[k], end_idx[k])`` returns (sswpy.pyx:227-304);
    ``gap_open`` / ``gap_extension`` / ``start_idx`` / ``end_idx`` may be scalars or per-pair sequences.
    Reads and references are de-duplicated tables: a locus' window is uploaded once however many reads
    and gap-penalty grid points refer to it.  Returns ``list[Alignment]``.
    """
    n = len(pair_read)
    pr = np.asarray(pair_read, dtype=np.int32)
    pw = np.asarray(pair_ref, dtype=np.int32)
    if pw.shape[0] != n:
        raise ValueError("pair_read and pair_ref must have the same length")
    if n and (pr.min() < 0 or pr.max() >= len(reads) or pw.min() < 0 or pw.max() >= len(references)):
        raise IndexError("pair index out of range")
    rdata, roff, rlen = _blob(reads)
    wdata, woff, wlen = _blob(references)

    def per_pair(v, default):
        if v is None:
            return np.full(n, default, dtype=np.int64)
        a = np.asarray(v, dtype=np.int64)
        return np.full(n, int(a), dtype=np.int64) if a.ndim == 0 else a

    go = per_pair(gap_open, 3)
    ge = per_pair(gap_extension, 1)
    s0 = per_pair(start_idx, 0)
    e0 = per_pair(end_idx, 0)
    ref_length = wlen[pw].astype(np.int64) if n else np.zeros(0, np.int64)
    if n:
        if (s0 < 0).any() or (e0 < 0).any():
            raise ValueError("negative indexing not supported")
        bad = (e0 > ref_length) | (s0 > ref_length)
        if bad.any():
            k = int(np.nonzero(bad)[0][0])
            raise ValueError("start_idx: {} or end_idx: {} can't be greater than ref_length: {}".format(int(s0[k]), int(e0[k]), int(ref_length[k])))
    e_final = np.where(e0 == 0, ref_length, e0)
    search = (e_final - s0).astype(np.int32)
    a = aligner or _aligner(device)
    res, arena = a.align(
        rdata, roff, rlen, wdata, woff, wlen, pr, pw,
        (go & 0xFF).astype(np.uint8), (ge & 0xFF).astype(np.uint8),
        ref_beg=s0.astype(np.int32), ref_len=search, mask_len=None,
        mat=dna_score_matrix(match_score, mismatch_penalty), n=5, score_size=2, flag=1,
        seq_encoding=L.SWB_SEQ_ASCII,
    )
    if n and (res["status"] != 0).any():
        raise ValueError("Problem Running alignment, see stdout")
    # Alignment tuples exactly like sswpy.pyx:283-298; the CIGAR tokens of the whole batch are formatted in one vectorised pass
    # ("%d%s" per op, MAPSTR "MIDNSHP=X", ssw.h:171-190: op codes above 8 print as 'M')
    ops_chars = np.array(list(_OPS + "M" * 7))
    tok = np.char.add((arena >> 4).astype("U"), ops_chars[arena & 15]).tolist() if arena.shape[0] else []
    out = [Alignment("".join(tok[o : o + l]) if l > 0 else None, a, b, c, d, e, f)
           for o, l, a, b, c, d, e, f in zip(res["cigar_off"].tolist(), res["cigar_len"].tolist(), res["score1"].tolist(), res["score2"].tolist(),
                                             res["ref_begin1"].tolist(), res["ref_end1"].tolist(), res["read_begin1"].tolist(), res["read_end1"].tolist())]
    return out


# ---------------------------------------------------------------------------------------------
# prefetch: batch now, answer the per-call API from memory later
# ---------------------------------------------------------------------------------------------
# indelPost's control flow asks for one alignment at a time (retarget / grid_search / is_target_by_ssw), but what it
# can ask for is known as soon as a locus' pileup is built: its reads x its windows (reference window, contig) x the
# gap-penalty grid (varaln.pyx:1127-1143).  prefetch_alignments() computes that whole set in ONE GPU batch and
# SSW.align() answers from it, so the unmodified per-call code runs at batch throughput.  Results are the same
# tuples align() would compute; nothing is approximated.

INDELPOST_GRID = ((3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0))          # varaln.pyx:1127-1143
_PREFETCHED: dict = {}
_PREFETCH_LIMIT = 4_000_000


def clear_prefetched():
    _PREFETCHED.clear()


def prefetch_alignments(reads: Sequence[STR_T], references: Sequence[STR_T], pair_read=None, pair_ref=None, grid=INDELPOST_GRID,
                        match_score: int = 3, mismatch_penalty: int = 2, device: int = 0) -> int:
    """Align every (read, reference) pair -- all combinations if ``pair_read`` / ``pair_ref`` are None -- under every
    (gap_open, gap_extension) of ``grid`` in one GPU batch and keep the results for ``SSW.align``.  A grid entry whose
    gap_open is the string ``"len"`` stands for ``gap_open = len(read)`` (localn.pyx:253-255).  Returns the number of
    alignments computed."""
    if pair_read is None or pair_ref is None:
        pair_read = np.repeat(np.arange(len(reads), dtype=np.int32), len(references))
        pair_ref = np.tile(np.arange(len(references), dtype=np.int32), len(reads))
    pr = np.asarray(pair_read, dtype=np.int32)
    pw = np.asarray(pair_ref, dtype=np.int32)
    n, g = pr.shape[0], len(grid)
    if n == 0 or g == 0:
        return 0
    raws_r = [_to_bytes(s) for s in reads]
    rlen = np.fromiter((len(r) for r in raws_r), dtype=np.int64, count=len(raws_r))
    go = np.empty((g, n), dtype=np.int64)
    ge = np.empty((g, n), dtype=np.int64)
    for k, (o, e) in enumerate(grid):
        go[k] = rlen[pr] if o == "len" else int(o)
        ge[k] = rlen[pr] if e == "len" else int(e)
    alns = align_batch(reads, references, np.tile(pr, g), np.tile(pw, g), go.reshape(-1), ge.reshape(-1),
                       match_score=match_score, mismatch_penalty=mismatch_penalty, device=device)
    if len(_PREFETCHED) + len(alns) > _PREFETCH_LIMIT:
        _PREFETCHED.clear()
    mkey = dna_score_matrix(_c_int(match_score), _c_int(mismatch_penalty)).tobytes()
    ref_keys = [_LUT[np.frombuffer(_to_bytes(s), dtype=np.uint8)].tobytes() for s in references]
    read_keys = [_LUT[np.frombuffer(r, dtype=np.uint8)].tobytes() for r in raws_r]
    ref_len = [len(k) for k in ref_keys]
    gof, gef = go.reshape(-1), ge.reshape(-1)
    for k, a in enumerate(alns):
        p = k % n
        r, w = int(pr[p]), int(pw[p])
        _PREFETCHED[(mkey, ref_keys[w], read_keys[r], int(gof[k]) & 0xFF, int(gef[k]) & 0xFF, 0, ref_len[w])] = a
    return len(alns)


# ---------------------------------------------------------------------------------------------
# convenience wrappers kept for API parity (sswpy.pyx:339-396)
# ---------------------------------------------------------------------------------------------

def force_align(read: STR_T, reference: STR_T, force_overhang: bool = False, aligner: SSW = None) -> Alignment:
    a = SSW() if aligner is None else aligner
    a.setRead(read)
    a.setReference(reference)
    len_x = len(read)
    res = a.align(gap_open=len_x)
    if res.optimal_score < 4:
        raise ValueError("No solution found")
    if force_overhang:
        if res.reference_start != 0 or res.reference_end != len(reference) - 1:
            raise ValueError("Read does not align to one overhang")
    return res


def _str(s):
    return s.decode("utf8") if isinstance(s, bytes) else s


def format_force_align(read: STR_T, reference: STR_T, alignment: Alignment, do_print: bool = False):
    start_ref = alignment.reference_start
    start_read = alignment.read_start
    buffer_ref = ""
    buffer_read = ""
    if start_ref < start_read:
        buffer_ref = " " * (start_read - start_ref)
    else:
        buffer_read = " " * (start_ref - start_read)
    ref_out = buffer_ref + _str(reference)
    read_out = buffer_read + _str(read)
    if do_print:
        print(ref_out)
        print(read_out)
    return ref_out, read_out
