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

import os

from typing import NamedTuple, Optional, Sequence, Union

import numpy as np

from . import _lib as L
from . import _swbhost as H
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
_RESOLVER = None          # set by indelpost_b200.wave while a wave scheduler runs: (ssw, go8, ge8, start_idx, search_length) -> Alignment | None


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


_MATRICES: dict = {}      # (match, mismatch) -> (5x5 int8 matrix, its bytes): built once, shared by every SSW with those scores


class SSW:
    """Drop-in for reference ``sswpy.SSW`` (sswpy.pyx:99-337).

    indelPost builds a fresh aligner for almost every call (`make_aligner` + `align`, localn.pyx:464-472: thousands per locus), and
    under prefetch / the wave scheduler nearly every `align()` is answered from a batch computed earlier.  So everything that is
    only needed by the single-pair GPU path -- encoding the sequences, `ssw_init`, the profile -- is deferred until that path is
    actually taken; `setRead` / `setReference` just record the sequence."""

    __slots__ = ("_lib", "score_matrix", "_mkey", "_ms", "_mm", "_rid", "_wid", "_rkey", "_wkey", "read", "reference", "_read_arr", "_ref_arr",
                 "_profile", "read_length", "ref_length", "_memo", "__dict__")

    def __init__(self, match_score: int = 2, mismatch_penalty: int = 2):
        self._lib = None
        m = _MATRICES.get((match_score, mismatch_penalty))
        if m is None:
            mat = dna_score_matrix(_c_int(match_score), _c_int(mismatch_penalty))        # buildDNAScoreMatrix
            mat.setflags(write=False)
            m = _MATRICES[(match_score, mismatch_penalty)] = (mat, mat.tobytes())
        self.score_matrix, self._mkey = m
        self._ms, self._mm = match_score, mismatch_penalty
        self._rid = self._wid = self._rkey = self._wkey = None
        self.read = None
        self.reference = None
        self._read_arr = None
        self._ref_arr = None
        self._profile = None
        self.read_length = 0
        self.ref_length = 0
        # results of this aligner's recent calls: indelPost repeats alignments with identical inputs (update_read_info,
        # pileup.pyx:849, re-runs what retarget computed at pileup.pyx:647 -- SURVEY.md 8f item 1); a hit skips the GPU round trip
        self._memo = None

    def __del__(self):  # __dealloc__, sswpy.pyx:135-147
        try:
            if self._profile and self._lib is not None:
                self._lib.init_destroy(self._profile)
                self._profile = None
        except Exception:
            pass

    def _drop_profile(self):
        if self._profile and self._lib is not None:
            self._lib.init_destroy(self._profile)
        self._profile = None

    def setRead(self, read: STR_T):  # sswpy.pyx:149-178
        raw = read.encode("utf8") if type(read) is str else (read if type(read) is bytes else _to_bytes(read))
        if self._profile:
            self._drop_profile()
        self.read = read
        self._read_arr = None
        self._rid = _SEQ_IDS.get(raw)            # None: this read is in no prefetched block
        self._rkey = raw
        self.read_length = len(raw)

    def setReference(self, reference: STR_T):  # sswpy.pyx:180-197
        raw = reference.encode("utf8") if type(reference) is str else (reference if type(reference) is bytes else _to_bytes(reference))
        self._ref_arr = None
        self._wid = _SEQ_IDS.get(raw)
        self._wkey = raw
        self.reference = reference
        self.ref_length = len(raw)
        self._memo = None

    def _single_pair(self, gap_open, gap_extension, start_idx, search_length):
        """the per-call GPU path (ssw_init + ssw_align through the compatibility symbols): a one-pair round trip"""
        if self._lib is None:
            self._lib = L.load()
        if self._ref_arr is None:
            self._ref_arr = np.ascontiguousarray(_LUT[np.frombuffer(self._wkey, dtype=np.uint8)])
        if not self._profile:
            self._read_arr = np.ascontiguousarray(_LUT[np.frombuffer(self._rkey, dtype=np.uint8)])
            self._profile = self._lib.ssw_init(self._read_arr.ctypes.data, self.read_length, self.score_matrix.ctypes.data, 5, 2)
            if not self._profile:
                raise ValueError("Must set profile first")
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
        return out

    def align(self, gap_open: int = 3, gap_extension: int = 1, start_idx: int = 0, end_idx: int = 0) -> Alignment:
        """sswpy.pyx:227-304 (same checks, same order, same messages)"""
        if type(gap_open) is not int or type(gap_extension) is not int or not (-2147483648 <= gap_open < 2147483648) or not (-2147483648 <= gap_extension < 2147483648):
            gap_open = _c_int(gap_open, "gap_open")
            gap_extension = _c_int(gap_extension, "gap_extension")
        ref_length = self.ref_length
        if start_idx == 0 and end_idx == 0:
            search_length = ref_length
        else:
            start_idx = int(start_idx)
            end_idx = int(end_idx)
            if start_idx < 0 or end_idx < 0:
                raise ValueError("negative indexing not supported")
            if end_idx > ref_length or start_idx > ref_length:
                raise ValueError(
                    "start_idx: {} or end_idx: {} can't be greater than ref_length: {}".format(start_idx, end_idx, ref_length)
                )
            search_length = (ref_length if end_idx == 0 else end_idx) - start_idx
        if self.reference is None:
            raise ValueError("call setReference first")
        rkey = self._rkey
        if rkey is None:
            raise ValueError("Must set profile first")
        go8, ge8 = gap_open & 0xFF, gap_extension & 0xFF
        memo = self._memo
        key = (rkey, go8, ge8, start_idx, search_length)
        if memo is not None:
            hit = memo.get(key)
            if hit is not None:
                return hit
        if search_length == ref_length and start_idx == 0:
            if _BLOCKS:
                rid, wid = self._rid, self._wid
                if rid is None:
                    rid = self._rid = _SEQ_IDS.get(rkey)
                if wid is None:
                    wid = self._wid = _SEQ_IDS.get(self._wkey)
                if rid is not None and wid is not None:
                    blocks = _BLOCKS.get((self._mkey, wid))                                           # filled by prefetch_alignments()
                    if blocks:
                        rlen = self.read_length
                        for b in blocks:
                            hit = b.find(rid, wid, go8, ge8, rlen)
                            if hit is not None:
                                return hit
        if _RESOLVER is not None:
            hit = _RESOLVER(self, go8, ge8, start_idx, search_length)                                 # wave scheduler (wave.py)
            if hit is not None:
                return hit
        if AUTO_BATCH and search_length == ref_length and start_idx == 0:
            hit = _auto_get(self, go8, ge8)                                                            # implicit batching (below)
            if hit is not None:
                return hit
        out = self._single_pair(gap_open, gap_extension, start_idx, search_length)
        if memo is None:
            memo = self._memo = {}
        elif len(memo) >= 256:
            memo.clear()
        memo[key] = out
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


class AlignmentList(Sequence):
    """Results of one batch: a sequence of ``Alignment`` tuples backed by the C ABI's fixed-stride records
    (``records``, dtype RESULT_DTYPE = every s_align field, ssw.h:55-66) and the BAM-packed CIGAR arena
    (``cigar_arena``).  An ``Alignment`` (and its "%d%s" CIGAR string, sswpy.pyx:283-298) is only built when an
    element is asked for; vectorised consumers read ``records`` directly."""

    __slots__ = ("records", "cigar_arena", "_lease", "_ra", "_aa")

    def __init__(self, records: np.ndarray, cigar_arena: np.ndarray, lease=None):
        self.records = records
        self.cigar_arena = cigar_arena
        self._lease = lease                       # keeps the pinned output buffers alive (returned to the aligner's pool on release)
        self._ra = records.ctypes.data
        self._aa = cigar_arena.ctypes.data

    def __len__(self):
        return self.records.shape[0]

    def __getitem__(self, k):
        n = self.records.shape[0]
        if isinstance(k, slice):
            return [H.alignment(self._ra, self._aa, i, Alignment) for i in range(*k.indices(n))]
        k = int(k)
        if k < 0:
            k += n
        if not 0 <= k < n:
            raise IndexError("alignment index out of range")
        return H.alignment(self._ra, self._aa, k, Alignment)

    def __iter__(self):
        n = self.records.shape[0]
        for k0 in range(0, n, 4096):
            yield from H.alignments(self._ra, self._aa, k0, min(n, k0 + 4096), Alignment)

    def tolist(self):
        return H.alignments(self._ra, self._aa, 0, self.records.shape[0], Alignment)

    def __eq__(self, other):
        if isinstance(other, AlignmentList):
            other = other.tolist()
        return self.tolist() == other

    __hash__ = None


class _Staging:
    """reusable pinned host arrays for the inputs of one batch (the batched entry's contract: SURVEY.md 8b).  Two sets per
    aligner, used alternately, so the arrays of the previous call stay untouched while a caller still looks at them."""

    def __init__(self):
        self.bufs = {}

    def view(self, name, dtype, count):
        dt = np.dtype(dtype)
        need = max(1, count) * dt.itemsize + 16
        b = self.bufs.get(name)
        if b is None or b.nbytes < need:
            if b is not None:
                b.close()
            b = self.bufs[name] = L.PinnedBuffer(need + need // 4)
        return b.view(dt, count)

    def table(self, name, seqs):
        """str / bytes sequences -> (blob, off, len) in pinned memory, copied once, in C (csrc/swbhost.c).  The blob keeps the
        capacity of earlier calls; if this table does not fit, gather reports the size it needs and is called again."""
        n = len(seqs)
        off = self.view(name + ".off", np.int64, n)
        ln = self.view(name + ".len", np.int32, n)
        b = self.bufs.get(name + ".blob")
        cap = b.nbytes - 16 if b is not None else 0
        if cap <= 0:
            cap = max(1 << 16, H.total_len(seqs))
        while True:
            blob = self.view(name + ".blob", np.int8, cap)
            end = H.gather(seqs, blob.ctypes.data, cap, off.ctypes.data, ln.ctypes.data, 0)
            if end >= 0:
                return blob[:end], off, ln
            cap = -end - 1


def _staging(a: BatchAligner) -> _Staging:
    sets = getattr(a, "_staging_sets", None)
    if sets is None:
        sets = a._staging_sets = [_Staging(), _Staging()]
        a._staging_turn = 0
    a._staging_turn ^= 1
    return sets[a._staging_turn]


def _per_pair(v, default, n, name):
    if v is None:
        return None if default is None else np.full(n, default, dtype=np.int64)
    a = np.asarray(v, dtype=np.int64)
    if a.ndim == 0:
        return np.full(n, int(a), dtype=np.int64)
    if a.shape != (n,):
        raise ValueError(f"{name} must be a scalar or have one entry per pair ({n}), got shape {a.shape}")
    return a


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
) -> AlignmentList:
    """Align ``reads[pair_read[k]]`` against ``references[pair_ref[k]]`` for every k in ONE GPU batch.

    Each result equals what ``SSW(match_score, mismatch_penalty)`` + ``setReference`` + ``setRead`` +
    ``align(gap_open[k], gap_extension[k], start_idx[k], end_idx[k])`` returns (sswpy.pyx:227-304);
    ``gap_open`` / ``gap_extension`` / ``start_idx`` / ``end_idx`` may be scalars or per-pair sequences.
    Reads and references are de-duplicated tables: a locus' window is uploaded once however many reads
    and gap-penalty grid points refer to it.  The sequences are gathered straight into reusable pinned staging
    arrays and the results come back as an ``AlignmentList`` (lazy ``Alignment`` tuples over the raw records).
    """
    pr = np.asarray(pair_read, dtype=np.int32)
    pw = np.asarray(pair_ref, dtype=np.int32)
    n = int(pr.shape[0])
    if pr.ndim != 1 or pw.shape != (n,):
        raise ValueError("pair_read and pair_ref must be one-dimensional and have the same length")
    if n and (pr.min() < 0 or pr.max() >= len(reads) or pw.min() < 0 or pw.max() >= len(references)):
        raise IndexError("pair index out of range")
    go = _per_pair(gap_open, 3, n, "gap_open")
    ge = _per_pair(gap_extension, 1, n, "gap_extension")
    s0 = _per_pair(start_idx, None, n, "start_idx")
    e0 = _per_pair(end_idx, None, n, "end_idx")
    a = aligner or _aligner(device)
    st = _staging(a)
    rdata, roff, rlen = st.table("reads", reads)
    wdata, woff, wlen = st.table("windows", references)
    p_pr = st.view("pair_read", np.int32, n); p_pr[:] = pr
    p_pw = st.view("pair_win", np.int32, n); p_pw[:] = pw
    p_go = st.view("gap_open", np.uint8, n); np.bitwise_and(go, 0xFF, out=p_go, casting="unsafe")
    p_ge = st.view("gap_ext", np.uint8, n); np.bitwise_and(ge, 0xFF, out=p_ge, casting="unsafe")
    p_rb = p_rl = None
    if s0 is not None or e0 is not None:
        ref_length = wlen[pw].astype(np.int64) if n else np.zeros(0, np.int64)
        s0 = np.zeros(n, np.int64) if s0 is None else s0
        e0 = np.zeros(n, np.int64) if e0 is None else e0
        if n:
            if (s0 < 0).any() or (e0 < 0).any():
                raise ValueError("negative indexing not supported")
            bad = (e0 > ref_length) | (s0 > ref_length)
            if bad.any():
                k = int(np.nonzero(bad)[0][0])
                raise ValueError("start_idx: {} or end_idx: {} can't be greater than ref_length: {}".format(int(s0[k]), int(e0[k]), int(ref_length[k])))
        p_rb = st.view("ref_beg", np.int32, n); p_rb[:] = s0
        p_rl = st.view("ref_len", np.int32, n); p_rl[:] = np.where(e0 == 0, ref_length, e0) - s0
    res, arena, lease = a.align_leased(
        rdata, roff, rlen, wdata, woff, wlen, p_pr, p_pw, p_go, p_ge, ref_beg=p_rb, ref_len=p_rl, mask_len=None,
        mat=dna_score_matrix(match_score, mismatch_penalty), n=5, score_size=2, flag=1, seq_encoding=L.SWB_SEQ_ASCII,
    )
    if n and (res["status"] != 0).any():
        raise ValueError("Problem Running alignment, see stdout")
    return AlignmentList(res, arena, lease)


# ---------------------------------------------------------------------------------------------
# prefetch: batch now, answer the per-call API from memory later
# ---------------------------------------------------------------------------------------------
# indelPost's control flow asks for one alignment at a time (retarget / grid_search / is_target_by_ssw), but what it
# can ask for is known as soon as a locus' pileup is built: its reads x its windows (reference window, contig) x the
# gap-penalty grid (varaln.pyx:1127-1143).  prefetch_alignments() computes that whole set in ONE GPU batch and
# SSW.align() answers from it, so the unmodified per-call code runs at batch throughput.  Results are the same
# tuples align() would compute; nothing is approximated.
#
# Storage: one _Block per prefetch call -- the raw result records of (pair, grid point) plus three small index dicts
# (read id -> row, window id -> column, penalty pair -> grid slot).  Registering a block costs O(reads + windows + grid),
# not O(pairs); an Alignment tuple is built when the control flow asks for that pair, and kept.

INDELPOST_GRID = ((3, 1), (3, 0), (5, 1), (5, 0), (4, 1), (4, 0))          # varaln.pyx:1127-1143
_PREFETCH_LIMIT = 16_000_000            # pairs kept before the oldest blocks are dropped
_SEQ_IDS: dict = {}                     # raw sequence bytes -> interned id
_BLOCKS: dict = {}                      # (matrix bytes, window id) -> [blocks holding that window]
_BLOCK_FIFO: dict = {}                  # id(block) -> block, oldest first
_prefetched_pairs = 0


_next_seq_id = 0


def seq_id(raw: bytes) -> int:
    """interned id of a sequence; ids are never reused (an SSW object may hold one across clear_prefetched())"""
    global _next_seq_id
    i = _SEQ_IDS.get(raw)
    if i is None:
        i = _SEQ_IDS[raw] = _next_seq_id
        _next_seq_id += 1
    return i


def _grid_slot(grid: dict, go8: int, ge8: int, rlen: int):
    """index of a (narrowed) penalty pair in a block's grid; "len" entries stand for len(read) (localn.pyx:253-255)"""
    g = grid.get((go8, ge8))
    if g is None and go8 == (rlen & 0xFF):
        g = grid.get(("len", ge8))
        if g is None and ge8 == (rlen & 0xFF):
            g = grid.get(("len", "len"))
    return g


class _Block:
    __slots__ = ("alist", "n", "rows", "cols", "n_cols", "pairs", "grid", "made", "mkey", "wids")

    def find(self, rid, wid, go8, ge8, rlen):
        if self.pairs is not None:
            p = self.pairs.get((rid, wid))
        else:
            r = self.rows.get(rid)
            c = self.cols.get(wid)
            p = None if r is None or c is None else r * self.n_cols + c
        if p is None:
            return None
        g = _grid_slot(self.grid, go8, ge8, rlen)
        if g is None:
            return None
        k = g * self.n + p
        hit = self.made.get(k)
        if hit is None:
            hit = self.made[k] = self.alist[k]
        return hit


def clear_prefetched():
    global _prefetched_pairs
    _BLOCKS.clear()
    _BLOCK_FIFO.clear()
    _SEQ_IDS.clear()
    _prefetched_pairs = 0
    clear_auto_batches()


def drop_block(b) -> None:
    """forget one prefetched block (lists are replaced, not edited: a concurrent reader keeps iterating its own copy)"""
    global _prefetched_pairs
    if _BLOCK_FIFO.pop(id(b), None) is None:
        return
    _prefetched_pairs -= len(b.alist)
    for w in b.wids:
        lst = _BLOCKS.get((b.mkey, w))
        if lst is not None:
            rest = [x for x in lst if x is not b]
            if rest:
                _BLOCKS[(b.mkey, w)] = rest
            else:
                _BLOCKS.pop((b.mkey, w), None)


def register_block(alist: AlignmentList, rids, wids, pair_read, pair_ref, grid, mkey: bytes, cross: bool):
    """make the results of a (pairs x grid) batch -- record k = g * n_pairs + p -- visible to SSW.align"""
    global _prefetched_pairs
    b = _Block()
    b.alist, b.mkey, b.made = alist, mkey, {}
    b.n = len(alist) // max(1, len(grid))
    if cross:
        b.rows = {r: i for i, r in enumerate(rids)}
        b.cols = {w: i for i, w in enumerate(wids)}
        b.n_cols, b.pairs = len(wids), None
    else:
        b.rows = b.cols = None
        b.n_cols = 0
        b.pairs = {(rids[r], wids[w]): p for p, (r, w) in enumerate(zip(pair_read.tolist(), pair_ref.tolist()))}
    b.grid = {}
    for g, (o, e) in enumerate(grid):
        b.grid.setdefault((o if o == "len" else int(o) & 0xFF, e if e == "len" else int(e) & 0xFF), g)
    b.wids = list(dict.fromkeys(wids))
    for w in b.wids:
        _BLOCKS[(mkey, w)] = _BLOCKS.get((mkey, w), []) + [b]
    _BLOCK_FIFO[id(b)] = b
    _prefetched_pairs += len(alist)
    while _prefetched_pairs > _PREFETCH_LIMIT and len(_BLOCK_FIFO) > 1:
        drop_block(next(iter(_BLOCK_FIFO.values())))
    return b


def prefetched(mkey: bytes, rid, wid, go8: int, ge8: int, rlen: int):
    """the prefetched Alignment of (read id, window id, narrowed penalties) or None"""
    blocks = _BLOCKS.get((mkey, wid))
    if blocks:
        for b in blocks:
            hit = b.find(rid, wid, go8, ge8, rlen)
            if hit is not None:
                return hit
    return None


def grid_penalties(grid, rlen_of_pair: np.ndarray):
    """(gap_open, gap_extension) arrays of shape [len(grid), n_pairs]; "len" stands for len(read) (localn.pyx:253-255)"""
    n, g = rlen_of_pair.shape[0], len(grid)
    go = np.empty((g, n), dtype=np.int64)
    ge = np.empty((g, n), dtype=np.int64)
    for k, (o, e) in enumerate(grid):
        go[k] = rlen_of_pair if o == "len" else int(o)
        ge[k] = rlen_of_pair if e == "len" else int(e)
    return go, ge


def prefetch_alignments(reads: Sequence[STR_T], references: Sequence[STR_T], pair_read=None, pair_ref=None, grid=INDELPOST_GRID,
                        match_score: int = 3, mismatch_penalty: int = 2, device: int = 0) -> int:
    """Align every (read, reference) pair -- all combinations if ``pair_read`` / ``pair_ref`` are None -- under every
    (gap_open, gap_extension) of ``grid`` in one GPU batch and keep the results for ``SSW.align``.  A grid entry whose
    gap_open is the string ``"len"`` stands for ``gap_open = len(read)`` (localn.pyx:253-255).  Returns the number of
    alignments computed."""
    cross = pair_read is None or pair_ref is None
    if cross:
        pr = np.repeat(np.arange(len(reads), dtype=np.int32), len(references))
        pw = np.tile(np.arange(len(references), dtype=np.int32), len(reads))
    else:
        pr = np.asarray(pair_read, dtype=np.int32)
        pw = np.asarray(pair_ref, dtype=np.int32)
    n, g = pr.shape[0], len(grid)
    if n == 0 or g == 0:
        return 0
    raws_r = [_to_bytes(s) for s in reads]
    raws_w = [_to_bytes(s) for s in references]
    rlen = np.fromiter(map(len, raws_r), dtype=np.int64, count=len(raws_r))
    go, ge = grid_penalties(grid, rlen[pr])
    alist = align_batch(raws_r, raws_w, np.tile(pr, g), np.tile(pw, g), go.reshape(-1), ge.reshape(-1),
                        match_score=match_score, mismatch_penalty=mismatch_penalty, device=device)
    mkey = dna_score_matrix(_c_int(match_score), _c_int(mismatch_penalty)).tobytes()
    register_block(alist, [seq_id(r) for r in raws_r], [seq_id(w) for w in raws_w], pr, pw, grid, mkey, cross)
    return len(alist)


# ---------------------------------------------------------------------------------------------
# implicit batching: what a ZERO-CHANGE caller gets (no prefetch line, no wave scheduler)
# ---------------------------------------------------------------------------------------------
# A single alignment through the GPU is a ~0.5 ms round trip against ~30 us for ssw.c, and it costs the same whether it carries
# one pair or ten thousand.  So an `SSW.align()` nobody prepared for does not fetch one alignment: it fetches
#   * every point of the gap grid indelPost can ask about this (read, window) (varaln.pyx:1127-1143), so grid_search's other five
#     calls are answered from memory;
#   * the same for the reads this process aligned most recently against OTHER windows: indelPost walks the same pileup over one
#     window after another (reference window, contig, retargeted windows), so the first call on a new window
#     brings the whole pileup along;
#   * a first `gap_open = len(read)` call on a window (is_target_by_ssw, localn.pyx:255; is_perfect_match, varaln.pyx:1230)
#     fetches those penalties for every read the window has been asked about -- in a batch of its own, because they run through
#     the exact kernels, whose latency the grid batches should not pay.
# What is computed is what the per-call path would compute (same kernels, same records); the cache is keyed by the sequences
# themselves and spans aligner objects (indelPost builds a new SSW for almost every call), which also serves the repeats of
# update_read_info (pileup.pyx:849).  The first window of a locus still costs one round trip per read: for batch throughput use
# prefetch_alignments() or the wave scheduler.  SWB200_AUTO_BATCH=0 turns this off (one pair per call).

AUTO_BATCH = os.environ.get("SWB200_AUTO_BATCH", "1") != "0"
_AUTO_GRID = tuple(INDELPOST_GRID)                              # what a first call on a (read, window) brings along
_AUTO_LEN_GRID = (("len", 1), ("len", 0), ("len", "len"))      # ... and a first `gap_open = len(read)` call (these take the exact kernels:
                                                                #     a millisecond of latency per batch, so they travel separately)
_AUTO_MAX_READS = 1024          # reads carried along per call ...
_AUTO_MAX_BASES = 1 << 20       # ... and their total length
_AUTO_RECENT = 2048             # distinct reads remembered per substitution matrix
_AUTO_LIMIT = 4_000_000         # alignments kept before the cache is dropped
_AUTO: dict = {}                # (matrix bytes, window bytes) -> {read bytes: [(AlignmentList, row, n_rows, grid map, made), ...]}
_RECENT: dict = {}              # matrix bytes -> {read bytes: None}, oldest first
_auto_pairs = 0
auto_stats = {"batches": 0, "pairs": 0, "hits": 0}


def clear_auto_batches():
    global _auto_pairs
    _AUTO.clear()
    _RECENT.clear()
    _auto_pairs = 0


def _auto_hit(entries, go8, ge8, rlen):
    for alist, row, n_rows, gmap, made, _tag in entries:
        g = _grid_slot(gmap, go8, ge8, rlen)
        if g is not None:
            k = g * n_rows + row
            hit = made.get(k)
            if hit is None:
                hit = made[k] = alist[k]
            return hit
    return None


def _grid_covers(grid, go8, ge8, rlen8):
    return any((o == go8 or (o == "len" and go8 == rlen8)) and (e == ge8 or (e == "len" and ge8 == rlen8)) for o, e in grid)


def _auto_get(ssw: "SSW", go8: int, ge8: int):
    """answer from the implicit cache, or compute this pair together with what is likely to be asked next; None = use the
    single-pair path (the batch could not be run: its error is the single-pair path's to report)"""
    global _auto_pairs
    mkey, wkey, rkey, rlen = ssw._mkey, ssw._wkey, ssw._rkey, ssw.read_length
    win = _AUTO.get((mkey, wkey))
    if win is not None:
        entries = win.get(rkey)
        if entries is not None:
            hit = _auto_hit(entries, go8, ge8, rlen)
            if hit is not None:
                auto_stats["hits"] += 1
                return hit
    else:
        win = _AUTO[(mkey, wkey)] = {}
    recent = _RECENT.get(mkey)
    if recent is None:
        recent = _RECENT[mkey] = {}
    rlen8 = rlen & 0xFF
    reads, bases = [rkey], rlen
    if _grid_covers(_AUTO_GRID, go8, ge8, rlen8):
        # the first question about this (read, window): the whole grid, and the recent reads this window has not seen
        grid, tag = _AUTO_GRID, 0
        candidates = (r for r in reversed(recent) if r not in win)
    elif _grid_covers(_AUTO_LEN_GRID, go8, ge8, rlen8):
        # the first `gap_open = len(read)` question (is_target_by_ssw, is_perfect_match): those penalties for every read this
        # window has been asked about
        grid, tag = _AUTO_LEN_GRID, 1
        candidates = (r for r, es in reversed(win.items()) if not any(e[5] == 1 for e in es))
    else:
        grid, tag = ((go8, ge8),), 2
        candidates = ()
    if wkey and rkey:
        for r in candidates:
            if r != rkey:
                reads.append(r)
                bases += len(r)
                if len(reads) >= _AUTO_MAX_READS or bases >= _AUTO_MAX_BASES:
                    break
    recent.pop(rkey, None)
    recent[rkey] = None
    if len(recent) > _AUTO_RECENT:
        del recent[next(iter(recent))]
    alist = None
    for attempt in (reads, [rkey]):
        try:
            n = len(attempt)
            rl = np.fromiter(map(len, attempt), dtype=np.int64, count=n)
            go, ge = grid_penalties(grid, rl)
            got = align_batch(attempt, [wkey], np.tile(np.arange(n, dtype=np.int32), len(grid)), np.zeros(n * len(grid), np.int32),
                              go.reshape(-1), ge.reshape(-1), match_score=ssw._ms, mismatch_penalty=ssw._mm)
            alist = AlignmentList(got.records.copy(), got.cigar_arena.copy())       # out of the pinned output buffers
            reads = attempt
            break
        except (ValueError, L.SwbError):
            if len(attempt) == 1 or len(reads) == 1:
                return None                      # e.g. an empty sequence, or no device: the single-pair path raises the reference's error
    if alist is None:
        return None
    gmap = {}
    for g, (o, e) in enumerate(grid):
        gmap.setdefault((o if o == "len" else int(o) & 0xFF, e if e == "len" else int(e) & 0xFF), g)
    made: dict = {}
    n = len(reads)
    for row, r in enumerate(reads):
        win.setdefault(r, []).append((alist, row, n, gmap, made, tag))
    auto_stats["batches"] += 1
    auto_stats["pairs"] += len(alist)
    _auto_pairs += len(alist)
    hit = _auto_hit(win[rkey], go8, ge8, rlen)
    if _auto_pairs > _AUTO_LIMIT:
        _AUTO.clear()
        _auto_pairs = 0
    return hit


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
