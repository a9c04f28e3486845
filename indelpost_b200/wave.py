"""Wave-batched host pipeline.

indelPost's control flow asks for one Smith-Waterman alignment at a time, and what it asks for next depends on the
last answer (grid winner -> new target -> contig -> localn; SURVEY.md §3.1: 3-5 sequential "waves" per locus,
varaln.pyx:1148-1225, pileup.pyx:577-808, localn.pyx:15-68).  A GPU wants all of it at once.  `WaveRunner` reconciles the
two WITHOUT touching that control flow:

  * every locus runs the unmodified per-call code (`VariantAlignment(...)`, `count_alleles`, `phase`, or anything else built
    on `make_aligner` / `align`) as a cooperative task -- a host thread that only ever runs between two alignment requests;
  * an `SSW.align()` that is not already answered from a prefetched block parks its task and files a REQUEST; when every
    live task is parked, the runner merges all requests -- many loci, many reads, the whole gap-penalty grid -- into ONE
    `swb_align_batch` call, registers the results as prefetch blocks and resumes the tasks;
  * a request is widened speculatively to what the locus will ask next: (every read of the locus) x (the window just
    asked about) x (the gap grid of varaln.pyx:1127-1143 + the `gap_open = len(read)` variants of localn.pyx:255 and
    varaln.pyx:1230), so `retarget`'s per-read loop, the six grid points, `update_read_info`'s repeats (pileup.pyx:849) and
    `is_target_by_ssw` are all answered by the wave that served the first miss.

Results are the tuples the per-call path would have produced (same kernels, same records); the control flow, and therefore
`count_alleles` / `phase`, cannot tell the difference (tests/test_pipeline_parity.py).  There is no CPU alignment anywhere:
a task whose request cannot be served raises.
"""
from __future__ import annotations

import threading
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

from . import sswpy
from .batch import dna_score_matrix
from .sswpy import INDELPOST_GRID, _to_bytes

_local = threading.local()

# what a locus can ask about a (read, window) pair: the grid + is_target_by_ssw's forced-gapless aligner with the extension
# penalty of any grid point + is_perfect_match (varaln.pyx:1127-1143, localn.pyx:255, varaln.pyx:1230)
DEFAULT_GRID = tuple(INDELPOST_GRID) + (("len", 1), ("len", 0), ("len", "len"))


class _Task:
    __slots__ = ("index", "reads", "read_set", "spec_done", "event", "error", "waves", "requests", "blocks")

    def __init__(self, index):
        self.index = index
        self.reads: List[bytes] = []
        self.read_set = set()
        self.spec_done = set()
        self.event = threading.Event()
        self.error = None
        self.waves = 0
        self.requests = 0
        self.blocks = []


class _Request:
    __slots__ = ("task", "mkey", "ms", "mm", "reads", "window", "grid")


def register_reads(seqs: Iterable) -> None:
    """tell the running wave task which reads its locus holds (call it from inside the function given to WaveRunner.map,
    e.g. right after the pileup is fetched); requests of this task are widened to all of them"""
    t = getattr(_local, "task", None)
    if t is None:
        return
    for s in seqs:
        b = _to_bytes(s)
        if b not in t.read_set:
            t.read_set.add(b)
            t.reads.append(b)


def tee_alignment_file(bam_cls):
    """subclass of a pysam-style AlignmentFile whose fetch() registers every read it yields with the running wave task"""

    class WaveAlignmentFile(bam_cls):
        def fetch(self, *a, **k):
            for r in super().fetch(*a, **k):
                seq = r.query_sequence
                if seq:
                    register_reads((seq,))
                yield r

    return WaveAlignmentFile


class WaveRunner:
    def __init__(self, device: int = 0, grid: Sequence = DEFAULT_GRID, max_inflight: int = 256, speculate: bool = True,
                 max_spec_windows: int = 16, aligner=None):
        self.device = device
        self.aligner = aligner or sswpy._aligner(device)
        self.grid = tuple(grid)
        self.max_inflight = max(1, int(max_inflight))
        self.speculate = speculate
        self.max_spec_windows = max_spec_windows
        self._mu = threading.Condition()
        # the baton: a task holds it whenever it runs and gives it up only while parked on a request, so exactly one task
        # executes host code at any time and tasks interleave ONLY at alignment requests -- code that touches process-wide
        # state between two requests (the reference seeds the global RNG before it downsamples, pileup.pyx:86-98) behaves as
        # it does single-threaded
        self._baton = threading.Lock()
        self._live = 0
        self._parked = 0
        self._pending: List[_Request] = []
        self.stats = {"waves": 0, "pairs": 0, "requests": 0, "tasks": 0}

    # ---- worker side -------------------------------------------------------------------------------------------------
    def _resolve(self, ssw, go8, ge8, start_idx, search_length):
        t = getattr(_local, "task", None)
        if t is None or getattr(_local, "runner", None) is not self:
            return None                                  # not one of our tasks: the per-call GPU path serves it
        if start_idx != 0 or search_length != ssw.ref_length:
            return None                                  # sub-range searches are not batched (indelPost never uses them)
        rkey, wkey = ssw._rkey, ssw._wkey
        rq = _Request()
        rq.task, rq.mkey = t, ssw._mkey
        rq.ms, rq.mm = ssw._ms, ssw._mm
        rq.window = wkey
        if rkey not in t.read_set:
            t.read_set.add(rkey)
            t.reads.append(rkey)
        if self.speculate and wkey not in t.spec_done and len(t.spec_done) < self.max_spec_windows:
            rq.reads = list(t.reads)
        else:
            rq.reads = [rkey]
        t.spec_done.add(wkey)
        rlen8 = ssw.read_length & 0xFF
        covered = any((o == go8 or (o == "len" and go8 == rlen8)) and (e == ge8 or (e == "len" and ge8 == rlen8)) for o, e in self.grid)
        rq.grid = self.grid if covered else self.grid + ((go8, ge8),)
        t.requests += 1
        with self._mu:
            self._pending.append(rq)
            self._parked += 1
            if self._parked >= self._live:
                self._mu.notify_all()
        self._baton.release()
        t.event.wait()
        t.event.clear()
        self._baton.acquire()
        if t.error is not None:
            raise t.error
        t.waves += 1
        ssw._rid = sswpy._SEQ_IDS.get(rkey)
        ssw._wid = sswpy._SEQ_IDS.get(wkey)
        hit = sswpy.prefetched(ssw._mkey, ssw._rid, ssw._wid, go8, ge8, ssw.read_length)
        if hit is None:
            raise RuntimeError("wave scheduler: a served request is missing from the prefetched blocks")
        return hit

    def _run_task(self, fn, item, task, results, reads):
        _local.task, _local.runner = task, self
        self._baton.acquire()
        try:
            if reads is not None:
                register_reads(reads(item))
            results[task.index] = (True, fn(item))
        except BaseException as e:  # noqa: BLE001 - handed to the caller of map()
            results[task.index] = (False, e)
        finally:
            _local.task = _local.runner = None
            for b in task.blocks:                        # the locus is done: nobody will ask for its alignments again
                sswpy.drop_block(b)
            task.blocks = []
            self._baton.release()
            with self._mu:
                self._live -= 1
                self._mu.notify_all()

    # ---- scheduler side ----------------------------------------------------------------------------------------------
    def _flush(self, reqs: List[_Request]):
        """all requests of a wave -> one batch per substitution matrix"""
        if len(sswpy._SEQ_IDS) > 2_000_000:
            sswpy._SEQ_IDS.clear()                       # ids are never reused: forgetting them only costs re-requests
        by_m = {}
        for rq in reqs:
            by_m.setdefault(rq.mkey, []).append(rq)
        for mkey, group in by_m.items():
            rtab, wtab = {}, {}
            pr_parts, pw_parts, go_parts, ge_parts, spans = [], [], [], [], []
            total = 0
            for rq in group:
                ridx = np.fromiter((rtab.setdefault(r, len(rtab)) for r in rq.reads), dtype=np.int32, count=len(rq.reads))
                w = wtab.setdefault(rq.window, len(wtab))
                n, g = ridx.shape[0], len(rq.grid)
                rl = np.fromiter(map(len, rq.reads), dtype=np.int64, count=n)
                go, ge = sswpy.grid_penalties(rq.grid, rl)
                pr_parts.append(np.tile(ridx, g))
                pw_parts.append(np.full(n * g, w, dtype=np.int32))
                go_parts.append(go.reshape(-1))
                ge_parts.append(ge.reshape(-1))
                spans.append((total, n * g))
                total += n * g
            alist = sswpy.align_batch(list(rtab), list(wtab), np.concatenate(pr_parts), np.concatenate(pw_parts),
                                      np.concatenate(go_parts), np.concatenate(ge_parts), match_score=group[0].ms, mismatch_penalty=group[0].mm,
                                      aligner=self.aligner)
            # out of the pinned output buffers (they go straight back to the pool); each request gets its slice of the records
            res = alist.records.copy()
            arena = alist.cigar_arena.copy()
            del alist
            for rq, (base, cnt) in zip(group, spans):
                sub = sswpy.AlignmentList(res[base: base + cnt], arena)
                rids = [sswpy.seq_id(r) for r in rq.reads]
                rq.task.blocks.append(sswpy.register_block(sub, rids, [sswpy.seq_id(rq.window)], None, None, rq.grid, mkey, cross=True))
            self.stats["pairs"] += total
        self.stats["waves"] += 1
        self.stats["requests"] += len(reqs)

    def map(self, fn: Callable, items: Sequence, reads: Optional[Callable] = None) -> list:
        """run fn(item) for every item as wave tasks; returns the results in order (an exception raised by a task is re-raised).
        `reads(item)`, if given, lists the read sequences of the item's locus (they widen its requests); alternatively the
        task calls wave.register_reads() itself, or fetches its reads through tee_alignment_file()."""
        items = list(items)
        results = [None] * len(items)
        prev = sswpy._RESOLVER
        sswpy._RESOLVER = self._resolve
        nxt = 0
        threads = []
        try:
            while True:
                with self._mu:
                    while nxt < len(items) and self._live < self.max_inflight:
                        task = _Task(nxt)
                        th = threading.Thread(target=self._run_task, args=(fn, items[nxt], task, results, reads), daemon=True)
                        self._live += 1
                        nxt += 1
                        threads.append(th)
                        th.start()
                    while self._live > 0 and self._parked < self._live:
                        self._mu.wait()
                    if self._live == 0 and nxt >= len(items):
                        break
                    reqs, self._pending = self._pending, []
                if reqs:
                    err = None
                    try:
                        self._flush(reqs)
                    except BaseException as e:  # noqa: BLE001 - every parked task must be released, with the error
                        err = e
                    with self._mu:
                        self._parked -= len(reqs)
                    for rq in reqs:
                        rq.task.error = err
                        rq.task.event.set()
        finally:
            sswpy._RESOLVER = prev
        for th in threads:
            th.join()
        self.stats["tasks"] += len(items)
        out = []
        for ok, v in results:
            if not ok:
                raise v
            out.append(v)
        return out
