"""Locus-parallel driver (SURVEY.md §8e "by locus for the VariantAlignment configs", §8f item 4 "keep in Python, parallelise
over loci").

Everything of a locus that is not Smith-Waterman -- pileup parsing, consensus, contig assembly, phasing (consensus.py:18-83,
contig.pyx:22-139, alleles.py:12-108) -- is the reference's own single-threaded Python/Cython and stays that way; loci are
independent (docs/benchmark.rst:11-13 recommends chunking), so the host side scales by PROCESSES and the alignment side by
the wave scheduler inside each process:

    pool = LocusPool(run_locus, workers=16, devices=range(8))      # run_locus: a picklable function of one work item
    results = pool.map(items)                                       # in item order
    pool.close()

  * `workers` host processes (spawned once, reused by every map()), worker k bound to devices[k % len(devices)]: with more
    workers than GPUs several processes share a device -- their waves interleave on it, which is what keeps a GPU busy while
    each process spends most of its time in the reference's Python;
  * inside a worker the items run as cooperative wave tasks (wave.WaveRunner): one merged `swb_align_batch` per wave over all
    loci the worker has in flight;
  * items are handed out in chunks from one shared queue (a worker that drew cheap loci simply draws again), results return
    through a second queue and are put back in order; a failing item re-raises in the caller with the worker's traceback.

There is no collective and no shared state between workers: results are per locus.  `mode="plain"` runs the
items without the wave scheduler (per-call `SSW.align`, or whatever the function uses) -- the arm the reference itself would
run under a process pool.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import traceback
from typing import Callable, Iterable, Optional, Sequence


def _worker_main(rank, device, fn, init, mode, max_inflight, tasks, results):
    try:
        os.environ.setdefault("SWB_BAM_THREADS", "2")          # many processes: keep the BGZF inflate pool small
        runner = None
        state = init(rank, device) if init is not None else None
        if mode == "wave":
            from . import wave

            runner = wave.WaveRunner(device=device, max_inflight=max_inflight, aligner=getattr(state, "aligner", None) if state is not None else None)
        results.put(("ready", rank, None, None))
    except BaseException:  # noqa: BLE001
        results.put(("fatal", rank, None, traceback.format_exc()))
        return
    while True:
        job = tasks.get()
        if job is None:
            break
        base, chunk = job
        try:
            if runner is not None:
                from . import sswpy

                out = runner.map(fn, chunk)
                sswpy.clear_prefetched()
                stats = dict(runner.stats)
            else:
                out = [fn(x) for x in chunk]
                stats = {}
            results.put(("ok", rank, base, (out, stats)))
        except BaseException:  # noqa: BLE001
            results.put(("error", rank, base, traceback.format_exc()))


class LocusPoolError(RuntimeError):
    pass


class LocusPool:
    def __init__(self, fn: Callable, workers: int = 1, devices: Sequence[int] = (0,), init: Optional[Callable] = None, mode: str = "wave",
                 max_inflight: int = 256, chunk: Optional[int] = None, start_timeout: float = 300.0):
        """fn(item) -> picklable result, a module-level function.  init(rank, device), optional, runs once in every worker
        before its first item (swap `indelpost.localn.SSW`, open files, ...); if it returns an object with an `aligner`
        attribute the worker's WaveRunner uses that aligner."""
        if mode not in ("wave", "plain"):
            raise ValueError("mode must be 'wave' or 'plain'")
        self.workers = max(1, int(workers))
        self.devices = list(devices) or [0]
        self.chunk = chunk
        self.mode = mode
        self.stats = {}
        ctx = mp.get_context("spawn")           # CUDA contexts do not survive fork
        self._tasks = ctx.Queue()
        self._results = ctx.Queue()
        self._procs = []
        for k in range(self.workers):
            p = ctx.Process(target=_worker_main, args=(k, self.devices[k % len(self.devices)], fn, init, mode, max_inflight, self._tasks, self._results), daemon=True)
            p.start()
            self._procs.append(p)
        import queue
        import time

        ready, deadline = 0, time.monotonic() + start_timeout
        while ready < self.workers:
            try:
                kind, rank, _, payload = self._results.get(timeout=1.0)
            except queue.Empty:
                dead = [k for k, p in enumerate(self._procs) if not p.is_alive()]
                if dead or time.monotonic() > deadline:
                    self.close()
                    raise LocusPoolError(f"worker(s) {dead} exited before they were ready" if dead else "a worker did not start in time") from None
                continue
            if kind == "fatal":
                self.close()
                raise LocusPoolError(f"worker {rank} failed to start:\n{payload}")
            ready += 1

    def map(self, items: Iterable) -> list:
        items = list(items)
        n = len(items)
        if n == 0:
            return []
        # chunks small enough to balance (about four per worker), large enough for a wave to merge many loci
        chunk = self.chunk or max(1, min(64, -(-n // (4 * self.workers))))
        jobs = 0
        for base in range(0, n, chunk):
            self._tasks.put((base, items[base: base + chunk]))
            jobs += 1
        out = [None] * n
        failure = None
        totals = {}
        for _ in range(jobs):
            kind, rank, base, payload = self._next_result()
            if kind == "ok":
                res, st = payload
                out[base: base + len(res)] = res
                for k, v in st.items():
                    totals[(rank, k)] = v            # runner.stats are cumulative per worker
            elif failure is None:
                failure = (rank, base, payload)
        self.stats = {}
        for (rank, k), v in totals.items():
            self.stats[k] = self.stats.get(k, 0) + v
        if failure is not None:
            raise LocusPoolError(f"worker {failure[0]} failed on the chunk starting at item {failure[1]}:\n{failure[2]}")
        return out

    def _next_result(self):
        """the next message of a worker; a worker that died without one (killed, out of memory, a crashed driver) must not leave
        the caller waiting forever"""
        import queue

        while True:
            try:
                return self._results.get(timeout=1.0)
            except queue.Empty:
                dead = [k for k, p in enumerate(self._procs) if not p.is_alive()]
                if dead:
                    codes = [self._procs[k].exitcode for k in dead]
                    self.close()
                    raise LocusPoolError(f"worker(s) {dead} exited without a result (exit codes {codes}); the pool is closed") from None

    def close(self):
        procs, self._procs = self._procs, []
        for _ in procs:
            try:
                self._tasks.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
