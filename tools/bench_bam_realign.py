"""Reads realigned per second FROM A BAM FILE (BASELINE.json's second metric on the pileup-shaped configs 1 / 3 / 4), end to end:

    BAM + BAI / FASTA + FAI on disk
      -> native columnar ingest per locus (indelpost_b200.bamio: region fetch, fetch_reads' filter, BAM 4-bit bases -> SWB_SEQ_PACKED4)
      -> every kept read x the locus' reference window (UnsplicedLocalReference.fetch_ref_seq, +-3 x window) x indelPost's six-point
         gap grid (varaln.pyx:1127-1143): the batch `grid_search` / `retarget` would issue one call at a time
      -> swb_align_batch on the GPU (loci merged into chunks; the ingest of the next chunk runs on host threads beside the GPU call)
      -> records + CIGARs back on the host.

The timed region starts with the files closed and ends with every record on the host.  Beside it: the reference's own ssw.c on
all host cores over a sample of the SAME pairs (alignment only -- pysam is absent, so its ingest cannot be timed at all).

    python tools/bench_bam_realign.py [--config cfg3|cfg4] [--loci N] [--threads T] [--chunk C]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID = ((3, 1), (3, 0), (4, 1), (4, 0), (5, 1), (5, 0))
SHAPES = {"cfg1": dict(n_reads=200, read_len=150, window=50, glen=4000),
          "cfg3": dict(n_reads=500, read_len=150, window=50, glen=4000),
          "cfg4": dict(n_reads=20000, read_len=250, window=167, glen=6000)}


def write_synthetic_bam(directory, n_loci, n_reads, read_len, glen, seed=1):
    """vectorised generator: every locus its own contig with one planted event (deletion or insertion of 1-20 bp, VAF 0.5, 0.5 %
    substitutions), reads as a mapper reports them (the gap in the CIGAR); written columnar by libswbbam.  -> (bam, fa, loci meta)"""
    from indelpost_b200 import bamio

    rng = np.random.default_rng(seed)
    L = read_len
    k = np.arange(L, dtype=np.int64)[None, :]
    acgt = np.frombuffer(b"ACGT", "u1")
    refs, seqs, meta = [], {}, []
    tid_l, pos_l, ncig_l, cig_l, seq_l, qual_l = [], [], [], [], [], []
    for t in range(n_loci):
        g = rng.integers(0, 4, glen)
        p = glen // 2                                   # 1-based position of the anchor base
        ev = int(rng.integers(1, 21))
        is_del = bool(t & 1)
        alt = rng.random(n_reads) < 0.5
        left = rng.integers(10, L - 10 - (0 if is_del else ev), n_reads)       # read bases up to and including the anchor
        start = np.sort(p - left)
        left = p - start
        ins = rng.integers(0, 4, ev)
        if is_del:
            src = start[:, None] + k + np.where(alt[:, None] & (k >= left[:, None]), ev, 0)
            r = g[np.clip(src, 0, glen - 1)]
        else:
            src = start[:, None] + k - np.where(alt[:, None] & (k >= left[:, None] + ev), ev, 0)
            r = g[np.clip(src, 0, glen - 1)]
            inside = alt[:, None] & (k >= left[:, None]) & (k < left[:, None] + ev)
            r = np.where(inside, ins[np.clip(k - left[:, None], 0, ev - 1)], r)
        sub = rng.random((n_reads, L)) < 0.005
        r = np.where(sub, (r + 1 + rng.integers(0, 3, (n_reads, L))) % 4, r)
        seq_l.append(acgt[r].reshape(-1))
        qual_l.append(rng.choice(np.array([30, 35, 37, 40], "u1"), (n_reads, L)).reshape(-1))
        right = L - left - (0 if is_del else ev)
        c = np.zeros((n_reads, 3), "<u4")
        c[:, 0] = np.where(alt, left << 4, L << 4)
        c[:, 1] = (ev << 4) | (2 if is_del else 1)
        c[:, 2] = right << 4
        nc = np.where(alt, 3, 1)
        cig_l.append(c[np.arange(3)[None, :] < nc[:, None]])
        ncig_l.append(nc)
        tid_l.append(np.full(n_reads, t)); pos_l.append(start)
        name = f"locus{t:06d}"
        refs.append((name, glen)); seqs[name] = acgt[g].tobytes().decode()
        meta.append((name, p))
    n = n_loci * n_reads
    n_cigar = np.concatenate(ncig_l)
    names = b"".join(b"r%08d\0" % i for i in range(n))
    bam_p, fa_p = os.path.join(directory, "loci.bam"), os.path.join(directory, "loci.fa")
    bamio.write_fasta(fa_p, seqs)
    bamio.write_bam_columns(bam_p, refs, np.concatenate(tid_l), np.concatenate(pos_l), np.zeros(n, "<u2"), np.full(n, 60, "u1"), np.full(n, L, "<i4"), n_cigar,
                            np.arange(n, dtype="<i8") * 10, np.arange(n, dtype="<i8") * L, np.concatenate([[0], np.cumsum(n_cigar)[:-1]]),
                            names, np.concatenate(seq_l).tobytes(), np.concatenate(qual_l).tobytes(), np.concatenate(cig_l), level=1)
    return bam_p, fa_p, meta


def ingest_locus(bam, fa, name, pos, window):
    """-> (packed read table, off, len, kept indices, window bytes) of one locus; C calls release the GIL"""
    from indelpost_b200 import bamio

    batch = bam.fetch_columns(name, max(0, pos - 1 - window), pos + window)
    keep = ((batch.flag & (bamio.FSECONDARY | bamio.FDUP)) == 0) & (batch.n_cigar > 0) & (batch.pos != 0)      # fetch_reads, pileup.pyx:138-147
    table, off, ln = batch.pack4()
    win = fa.fetch_bytes(name, max(0, pos - 3 * window), pos + 3 * window)                                       # local_reference.pyx:22-30
    return table, off, ln, keep.nonzero()[0], win


def build_chunk(items):
    """merge the loci of a chunk into one swb_align_batch call's arrays (PACKED4 tables)"""
    from indelpost_b200.batch import pack_table

    tables, offs, lens, pr, pw, wins = [], [], [], [], [], []
    base_read = base_byte = 0
    g = len(GRID)
    for w, (table, off, ln, keep, win) in enumerate(items):
        tables.append(table); offs.append(off + base_byte); lens.append(ln)
        pr.append(np.repeat(keep + base_read, g)); pw.append(np.full(keep.shape[0] * g, w, "<i4"))
        wins.append(win)
        base_read += off.shape[0]; base_byte += table.shape[0]
    pr = np.concatenate(pr).astype("<i4"); pw = np.concatenate(pw)
    n_kept = pr.shape[0] // g
    go = np.tile(np.array([a for a, _ in GRID], "u1"), n_kept); ge = np.tile(np.array([e for _, e in GRID], "u1"), n_kept)
    wlen = np.array([len(x) for x in wins], "<i4"); woff = np.concatenate([[0], np.cumsum(wlen[:-1])]).astype("<i8")
    wtab, wtoff = pack_table(np.frombuffer(b"".join(wins), "u1"), woff, wlen, bits=4, ascii=True)
    return dict(reads=np.concatenate(tables).view(np.int8), read_off=np.concatenate(offs), read_len=np.concatenate(lens), windows=wtab.view(np.int8), win_off=wtoff,
                win_len=wlen, pair_read=pr, pair_win=pw, gap_open=go, gap_ext=ge, n_reads_kept=n_kept)


def ingest_chunk_bulk(bam, fa, part, window, threads):
    """a whole chunk of loci in TWO C calls (swb_bam_fetch_pack4 on host threads + swb_fai_fetch_many) and a handful of numpy
    operations: no Python work per locus or per read"""
    from indelpost_b200 import bamio
    from indelpost_b200.batch import pack_table

    pk = bam.fetch_pack4([(name, max(0, pos - 1 - window), pos + window) for name, pos in part], exclude=bamio.FSECONDARY | bamio.FDUP, need_cigar=True,
                         drop_pos0=True, threads=threads)                                                        # fetch_reads, pileup.pyx:138-147
    wblob, woff = fa.fetch_many([(name, max(0, pos - 3 * window), pos + 3 * window) for name, pos in part])   # local_reference.pyx:22-30
    wlen = np.diff(woff).astype("<i4")
    wtab, wtoff = pack_table(wblob, woff[:-1], wlen, bits=4, ascii=True)
    g = len(GRID)
    pr = np.repeat(np.arange(pk.n_reads, dtype="<i4"), g)
    pw = np.repeat(pk.region_of_read(), g)
    go = np.tile(np.array([a for a, _ in GRID], "u1"), pk.n_reads); ge = np.tile(np.array([e for _, e in GRID], "u1"), pk.n_reads)
    return dict(reads=pk.table.view(np.int8), read_off=pk.read_off, read_len=pk.read_len, windows=wtab.view(np.int8), win_off=wtoff, win_len=wlen,
                pair_read=pr, pair_win=pw, gap_open=go, gap_ext=ge, n_reads_kept=pk.n_reads)


def measure(config="cfg3", n_loci=1000, threads=None, chunk=250, device=0, aligner=None, cpu_sample=24000, mode="bulk"):
    import swbtest as T
    from indelpost_b200 import bamio

    sh = SHAPES[config]
    threads = threads or max(2, min(16, len(os.sched_getaffinity(0))))
    tmp = tempfile.mkdtemp(prefix="swb_realign_")
    t0 = time.perf_counter()
    bam_p, fa_p, meta = write_synthetic_bam(tmp, n_loci, sh["n_reads"], sh["read_len"], sh["glen"])
    t_gen = time.perf_counter() - t0
    mat = T.dna_matrix(3, 2)
    if aligner is None:
        from indelpost_b200 import BatchAligner

        aligner = BatchAligner(device)
    window = sh["window"]

    def align_chunk(b, keep_results):
        return aligner.align(b["reads"], b["read_off"], b["read_len"], b["windows"], b["win_off"], b["win_len"], b["pair_read"], b["pair_win"],
                             b["gap_open"], b["gap_ext"], mat=mat, seq_encoding=2, copy=keep_results)

    def chunk_cells(b):       # bookkeeping for the GCUPS figure
        return int(np.dot(np.bincount(b["pair_win"], weights=b["read_len"][b["pair_read"]], minlength=b["win_len"].shape[0]), b["win_len"]))

    def run_bulk(keep_results=False):
        bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)
        chunks = [meta[i: i + chunk] for i in range(0, len(meta), chunk)]
        bg = ThreadPoolExecutor(max_workers=1)              # the next chunk's ingest (C calls, GIL released) beside this chunk's GPU call
        n_pairs = n_kept = cells = 0
        out = []
        pending = bg.submit(ingest_chunk_bulk, bam, fa, chunks[0], window, threads)
        for ci in range(len(chunks)):
            b = pending.result()
            if ci + 1 < len(chunks):
                pending = bg.submit(ingest_chunk_bulk, bam, fa, chunks[ci + 1], window, threads)
            res, arena = align_chunk(b, keep_results)
            n_pairs += res.shape[0]; n_kept += b["n_reads_kept"]; cells += chunk_cells(b)
            if keep_results:
                out.append((b, res, arena))
        bg.shutdown()
        bam.close(); fa.close()
        return n_pairs, n_kept, cells, out

    def run_per_locus(keep_results=False):
        pool = ThreadPoolExecutor(max_workers=threads)
        chunks = [meta[i: i + chunk] for i in range(0, len(meta), chunk)]
        # a reader handle is not thread-safe (block cache, span buffer) and cheap to open: one per ingest thread
        handles = [(bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)) for _ in range(threads)]

        def ingest_part(args):
            h, part = args
            return [ingest_locus(handles[h][0], handles[h][1], name, pos, window) for name, pos in part]

        def submit(ch):
            parts = [ch[i::threads] for i in range(threads)]
            return [pool.submit(ingest_part, (h, p)) for h, p in enumerate(parts) if p]

        n_pairs = n_kept = cells = 0
        out = []
        pending = submit(chunks[0])
        for ci in range(len(chunks)):
            items = [x for f in pending for x in f.result()]
            if ci + 1 < len(chunks):
                pending = submit(chunks[ci + 1])             # the next chunk's ingest runs beside this chunk's GPU call
            b = build_chunk(items)
            res, arena = align_chunk(b, keep_results)
            n_pairs += res.shape[0]; n_kept += b["n_reads_kept"]; cells += chunk_cells(b)
            if keep_results:
                out.append((b, res, arena))
        pool.shutdown()
        for hb, hf in handles:
            hb.close(); hf.close()
        return n_pairs, n_kept, cells, out

    run = run_bulk if mode == "bulk" else run_per_locus
    run()                                                    # warm-up: context, kernels, page cache
    t0 = time.perf_counter()
    n_pairs, n_kept, cells, _ = run()
    dt = time.perf_counter() - t0
    res = {"config": config, "loci": n_loci, "reads_in_bam": n_loci * sh["n_reads"], "reads_realigned": n_kept, "pairs": n_pairs, "grid_points": len(GRID),
           "read_len": sh["read_len"], "window_len": 6 * window, "host_threads": threads, "chunk_loci": chunk, "ingest": "one swb_bam_fetch_pack4 + one swb_fai_fetch_many call per chunk" if mode == "bulk" else "one fetch per locus on a Python thread pool", "bam_bytes": os.path.getsize(bam_p),
           "seconds": dt, "reads_per_s": n_kept / dt, "pairs_per_s": n_pairs / dt, "gcups": cells / dt / 1e9, "generate_and_write_s": t_gen,
           "what": "files closed -> every record and CIGAR on the host: native BAM ingest on host threads + swb_align_batch (PACKED4 tables straight from the BAM nibbles), "
                   "six gap-grid points per read against the locus window"}
    # the reference's ssw.c on a sample of the same pairs, all host cores (alignment only)
    try:
        sys.path.insert(0, ROOT)
        import bench as B

        bam, fa = bamio.AlignmentFile(bam_p), bamio.FastaFile(fa_p)
        items = [ingest_locus(bam, fa, name, pos, window) for name, pos in meta[: max(1, cpu_sample // (sh["n_reads"] * len(GRID)) + 1)]]
        b = build_chunk(items)
        n = min(cpu_sample, b["pair_read"].shape[0])
        # the CPU checker takes one code per byte: unpack the sample's tables
        def unpack(tab, off, ln):
            tab = tab.view(np.uint8)
            outl = []
            for o, l in zip(off, ln):
                t = tab[int(o): int(o) + (int(l) + 1) // 2]
                u = np.empty(2 * len(t), "i1"); u[0::2] = t & 15; u[1::2] = t >> 4
                outl.append(u[: int(l)])
            return outl
        cb = T.batch_from_lists(unpack(b["reads"], b["read_off"], b["read_len"]), unpack(b["windows"], b["win_off"], b["win_len"]),
                                b["pair_read"][:n], b["pair_win"][:n], b["gap_open"][:n], b["gap_ext"][:n], mat=mat)
        cores = B.host_cores()
        B.cpu_align_parallel(cb, cores)
        cdt, kind = B.cpu_align_parallel(cb, cores)
        res["cpu_baseline"] = {"kind": kind, "cores": cores, "pairs_per_s": n / cdt, "reads_per_s": n / cdt / len(GRID), "gcups": cb.cells() / cdt / 1e9,
                               "sample": f"{n} of the same pairs, alignment only (no ingest: pysam is absent)"}
        res["speedup_reads_per_s"] = res["reads_per_s"] / res["cpu_baseline"]["reads_per_s"]
    except Exception as e:  # noqa: BLE001
        res["cpu_baseline"] = {"error": repr(e)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3", choices=sorted(SHAPES))
    ap.add_argument("--loci", type=int, default=1000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=250)
    ap.add_argument("--mode", default="bulk", choices=("bulk", "per_locus"))
    a = ap.parse_args()
    print(json.dumps(measure(a.config, a.loci, a.threads or None, a.chunk, mode=a.mode)))


if __name__ == "__main__":
    main()
