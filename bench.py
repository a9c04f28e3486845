#!/usr/bin/env python
"""bench.py — SW GCUPS / reads realigned per second of the batched realignment hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]

Workload (BASELINE.json configs[1]): P = 1 M read x window pairs per GPU, 150-bp reads vs 400-bp
windows, score + coordinates + CIGAR (flag=1), (go, ge) = (3, 1), match 3 / mismatch 2, one planted
event per read (1/3 none, 1/3 deletion 1-10 bp, 1/3 insertion 1-10 bp) and 1 % substitutions, synthetic.
GCUPS := sum(readLen * windowLen) / seconds / 1e9 (one nominal forward matrix per pair).

A "step" is one pass of the whole hot path (prepare, forward, reverse, banded traceback) over the batch.
  value : inputs already resident in HBM when the timed region starts (swb_upload done; K x swb_compute)
  e2e   : the same through the one-shot C-ABI call swb_align_batch with pinned HOST buffers (sequence tables 2-bit packed,
          SWB_SEQ_PACKED2): H2D of every input + unpack + kernels + D2H of results and CIGARs inside the timed region;
          `e2e_codes` is the same call with one-byte-per-base tables (the round-1 figure), `extra.e2e_api` the Python entry
          `align_batch` from lists of ASCII sequences (host gather included)
  extra : (1 GPU only) the other BASELINE configs and mixes, each resident + e2e with its own stage split: cfg4 / cfg5 shapes,
          cfg2 with shared windows, indelPost's penalty mix, short reads; reads realigned per second from BAM + FASTA files
          (tools/bench_bam_realign.py); pileup ingestion rates (tools/bench_ingest.py); and the reference pipeline's loci/s on
          ssw.c vs under the wave scheduler, from memory and from BAM files, zero-change use, and eight workers through LocusPool
          (tools/bench_pipeline.py)
Multi-GPU (torchrun, one rank per GPU): the pairs shard across ranks with no collective in the data
path (weak scaling: every rank aligns its own P pairs); torch.distributed is only the barrier and the
max-over-ranks of the timed region.

--impl reference times the reference's own ssw.c (oracle/_ref, compiled from /root/reference by
oracle/Makefile) on the host cores with one thread per core on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "sw_gcups"
UNIT = "GCUPS"
READ_LEN, WIN_LEN = 150, 400


def workload_config(pairs, n_gpus):
    return {
        "workload": "cfg2: 1M read x window pairs per GPU, 150bp reads vs 400bp windows, score+coords+CIGAR (flag=1), go=3 ge=1, match=3 mismatch=2",
        "pairs_per_gpu": pairs,
        "read_len": READ_LEN,
        "window_len": WIN_LEN,
        "distinct_windows": True,
        "parallelism": f"pair-sharded x{n_gpus}, no collective",
        "l2_policy": "inputs (~550 MB/GPU unpacked) larger than the 126 MB L2; no explicit flush",
        "e2e_host_buffers": "pinned; sequence tables 2-bit packed (SWB_SEQ_PACKED2), unpacked on the device",
    }


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def _nvml_start(self):
        # in-process NVML polling (every ~4 ms): a 100 ms timed region still gets tens of samples
        import pynvml as nv
        nv.nvmlInit()
        h = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [x for x in vis.split(",") if x.strip().isdigit()]
            h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
        nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)          # fail here, not in the thread
        self.nv, self.h, self.samples, self.stop_flag = nv, h, [], False
        self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))

        def poll():
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                try:
                    self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                except Exception:
                    pass
                time.sleep(0.004)
        self.t = threading.Thread(target=poll, daemon=True)
        self.t.start()

    def _nvml_stop(self, t0, t1):
        self.stop_flag = True
        self.t.join(timeout=1.0)
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}   # nvml.h nvmlClocksEventReason*
        sm = [c for ts, c, r in self.samples if t0 <= ts <= t1]
        reasons = sorted({name for ts, c, r in self.samples if t0 <= ts <= t1 for name, b in bits.items() if r & b})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm), "source": "nvml"}

    def start(self):
        self.nvml = False
        try:
            self._nvml_start()
            self.nvml = True
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml:
            return self._nvml_stop(t0, t1)
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 <= ts <= t1 + 0.2:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.2:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU legs (the reference's own ssw.c when oracle/_ref is present, else the oracle port)
# ----------------------------------------------------------------------------------------------
def cpu_align_parallel(batch, n_threads):
    """time ref_align_batch / orc_align_batch over `batch`, pairs split evenly over n_threads threads
    (ctypes releases the GIL, each thread runs the C loop on its own core)."""
    import swbtest as T

    if T.have_ref():
        kind = "reference"
        chk = T.reference()
        parts = np.array_split(np.arange(batch.n_pairs), n_threads)
        subs = [(int(p[0]), int(p.shape[0])) for p in parts if p.shape[0]]
        run = lambda fc: chk.align_batch(batch, fc[0], fc[1])
    else:
        kind = "port"
        chk = T.oracle()
        parts = np.array_split(np.arange(batch.n_pairs), n_threads)
        subs = [batch.subset(p) for p in parts if p.shape[0]]
        run = lambda sb: chk.align_batch(sb)
    from concurrent.futures import ThreadPoolExecutor

    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)  # the reference prints warnings on stderr
    try:
        with ThreadPoolExecutor(max_workers=n_threads) as ex:
            t0 = time.perf_counter()
            list(ex.map(run, subs))
            dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)
    return dt, kind


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    import swbtest as T

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = host_cores()
    sample_pairs = max(cores * 250, min(args.pairs, cores * 1500))
    b = T.make_pairs_fast(sample_pairs, READ_LEN, WIN_LEN, seed=1234)
    for _ in range(args.warmup):
        cpu_align_parallel(b.subset(np.arange(min(sample_pairs, cores * 100))), cores)
    times = []
    kind = "reference"
    for _ in range(args.steps):
        dt, kind = cpu_align_parallel(b, cores)
        times.append(dt)
    tot = float(sum(times))
    gcups = b.cells() * args.steps / tot / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/s16 (SSE2)",
        "data": "synthetic", "config": workload_config(args.pairs, args.gpus),
        "reads_per_s": sample_pairs * args.steps / tot,
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_pairs} pairs of the cfg2 workload per step, split evenly over {cores} threads (unmodified ssw.c: ssw_init+ssw_align+destroy per pair)"},
        "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def dpx_peak():
    """measured DPX/ALU-pipe peak (profiles/dpx_peak.json, from tools/dpx_microbench on this pool's B200)"""
    p = os.path.join(ROOT, "profiles", "dpx_peak.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {"lanes_per_clk_per_sm": 64.0, "sm_mhz": 1965.0, "sms": 148, "instr_per_cell_pair": 6, "source": "fallback (nominal)"}


_ORIG_AFFINITY = None


def bind_to_gpu_numa(local):
    """multi-rank runs: keep this rank's threads and its pinned buffers (first touch) on the NUMA node its GPU hangs off, so the
    584 MB a step streams to the device do not cross the socket interconnect.  Best effort: silently skipped if sysfs says nothing."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        global _ORIG_AFFINITY
        _ORIG_AFFINITY = set(allowed)
        use = cpus & allowed
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            print(f"rank-local GPU {local} ({bdf}): bound to {len(use)} of {len(allowed)} cpus (GPU-local NUMA node)", file=sys.stderr)
    except Exception as e:  # noqa: BLE001
        print(f"NUMA binding skipped: {e}", file=sys.stderr)


def pinned_args(L, b, bits=0):
    """the batch's input arrays in pinned host memory (bits = 0: one code per byte; 2 / 4: packed tables)"""
    from indelpost_b200.batch import pack_table

    def pin(a):
        buf = L.PinnedBuffer(a.nbytes)
        v = buf.view(a.dtype, a.size)
        v[:] = a.reshape(-1)
        return buf, v

    src = {k: getattr(b, k) for k in ("reads", "read_off", "read_len", "windows", "win_off", "win_len", "pair_read", "pair_win", "gap_open", "gap_ext")}
    if bits:
        src["reads"], src["read_off"] = pack_table(b.reads, b.read_off, b.read_len, bits=bits)
        src["windows"], src["win_off"] = pack_table(b.windows, b.win_off, b.win_len, bits=bits)
        src["reads"] = src["reads"].view(np.int8)
        src["windows"] = src["windows"].view(np.int8)
    keep, arrs = [], []
    for k in ("reads", "read_off", "read_len", "windows", "win_off", "win_len", "pair_read", "pair_win", "gap_open", "gap_ext"):
        buf, v = pin(src[k])
        keep.append(buf)
        arrs.append(v)
    enc = {0: L.SWB_SEQ_CODES, 4: L.SWB_SEQ_PACKED4, 2: L.SWB_SEQ_PACKED2}[bits]
    return keep, arrs, enc


def measure_shape(al, L, b, steps=3, warmup=2, bits=2, peak_gcups=None):
    """resident and end-to-end GCUPS of one synthetic batch on one GPU, with the stage split of the resident leg"""
    kw = dict(mat=b.mat, n=5, score_size=2, flag=1)
    keep, arrs, enc = pinned_args(L, b, bits)
    keep0, arrs0, _ = pinned_args(L, b, 0)
    al.upload(*arrs0, **kw)
    for _ in range(warmup):
        al.compute()
    t0 = time.perf_counter()
    stage = {}
    for _ in range(steps):
        al.compute()
        tm = al.timing()
        for k in ("ms_prepare", "ms_forward", "ms_reverse", "ms_traceback", "ms_band_round0", "ms_band_rest", "ms_certify"):
            stage[k] = stage.get(k, 0.0) + tm[k] / steps
    dt = (time.perf_counter() - t0) / steps
    for _ in range(warmup):
        al.align(*arrs, copy=False, seq_encoding=enc, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        al.align(*arrs, copy=False, seq_encoding=enc, **kw)
    de = (time.perf_counter() - t0) / steps
    t = al.timing()
    cells = b.cells()
    out = {"pairs": int(b.n_pairs), "resident_gcups": cells / dt / 1e9, "e2e_gcups": cells / de / 1e9, "ms_resident": dt * 1e3, "ms_e2e": de * 1e3,
           "reads_per_s_resident": b.n_pairs / dt, "reads_per_s_e2e": b.n_pairs / de,
           "h2d_bytes": int(t["h2d_bytes"]), "d2h_bytes": int(t["d2h_bytes"]), "stage_ms": {k: round(v, 3) for k, v in stage.items()},
           "n_fast": int(tm["n_fast"]), "n_exact": int(tm["n_exact"])}
    if peak_gcups and stage.get("ms_forward", 0) > 0:
        out["forward_frac_of_dpx_peak"] = cells / (stage["ms_forward"] * 1e-3) / 1e9 / peak_gcups
    for bf in keep + keep0:
        bf.close()
    return out


def extra_measurements(al, L, T, peak_gcups):
    """the other BASELINE configs / mixes on one GPU (bounded sizes: the whole block takes about a minute)"""
    ex = {}
    shapes = [
        ("cfg2_shared_windows_5000x200", dict(n_pairs=1_000_000, read_len=150, win_len=400, seed=2, reads_per_window=200)),
        ("cfg4_250x1000", dict(n_pairs=120_000, read_len=250, win_len=1000, seed=4, max_indel=30)),
        ("cfg5_250x2000", dict(n_pairs=60_000, read_len=250, win_len=2000, seed=5, max_indel=10)),
        ("reads_100x300", dict(n_pairs=400_000, read_len=100, win_len=300, seed=3)),
        ("reads_75x300", dict(n_pairs=400_000, read_len=75, win_len=300, seed=6, max_indel=5)),
        ("reads_50x300", dict(n_pairs=400_000, read_len=50, win_len=300, seed=7, max_indel=3)),
    ]
    cfgs = {}
    for name, kw in shapes:
        try:
            cfgs[name] = measure_shape(al, L, T.make_pairs_fast(**kw), peak_gcups=peak_gcups)
        except Exception as e:  # noqa: BLE001
            cfgs[name] = {"error": repr(e)}
    ex["configs"] = cfgs
    # indelPost's own penalty mix (recorded call stream: ge = 0 on 40 % of the calls, go = len(read) on 6 %)
    try:
        n = 400_000
        b = T.make_pairs_fast(n, 150, 300, seed=11, reads_per_window=500)
        rng = np.random.default_rng(2)
        combos = np.array([(3, 1), (5, 1), (3, 0), (5, 0), (4, 1), (4, 0), (150, 1)], dtype=np.uint8)
        pick = rng.choice(7, size=n, p=[0.223, 0.18, 0.1344, 0.1344, 0.1344, 0.1344, 0.0594])
        b.gap_open = np.ascontiguousarray(combos[pick, 0])
        b.gap_ext = np.ascontiguousarray(combos[pick, 1])
        ex["grid_mix"] = dict(measure_shape(al, L, b, peak_gcups=peak_gcups), workload="indelPost penalty mix, 150 bp x 300 bp, 500 reads per window")
    except Exception as e:  # noqa: BLE001
        ex["grid_mix"] = {"error": repr(e)}
    # the Python entry point from lists of ASCII sequences: host gather + H2D + kernels + D2H, lazy result list
    try:
        from indelpost_b200 import align_batch

        n = 200_000
        b = T.make_pairs_fast(n, READ_LEN, WIN_LEN, seed=5)
        lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
        rs = [lut[b.reads[o:o + l]].tobytes() for o, l in zip(b.read_off.tolist(), b.read_len.tolist())]
        ws = [lut[b.windows[o:o + l]].tobytes() for o, l in zip(b.win_off.tolist(), b.win_len.tolist())]
        for _ in range(2):
            out = align_batch(rs, ws, b.pair_read, b.pair_win, 3, 1, match_score=3, mismatch_penalty=2, aligner=al)
        t0 = time.perf_counter()
        K = 3
        for _ in range(K):
            out = align_batch(rs, ws, b.pair_read, b.pair_win, 3, 1, match_score=3, mismatch_penalty=2, aligner=al)
        dt = (time.perf_counter() - t0) / K
        t1 = time.perf_counter()
        first = out[:1000]
        dmat = (time.perf_counter() - t1) / 1000
        ex["e2e_api"] = {"call": "indelpost_b200.align_batch(list[bytes], list[bytes], ...) -> AlignmentList", "pairs": n, "gcups": b.cells() / dt / 1e9,
                         "reads_per_s": n / dt, "ms": dt * 1e3, "us_per_materialised_alignment": dmat * 1e6, "sample_cigar": first[0].CIGAR}
    except Exception as e:  # noqa: BLE001
        ex["e2e_api"] = {"error": repr(e)}
    # reads realigned per second FROM A BAM FILE: native columnar ingest on host threads + the six-point gap grid on the GPU
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_bam_realign as BR

        ex["bam_realign"] = {"cfg3": BR.measure("cfg3", n_loci=600, chunk=200, aligner=al)}
    except Exception as e:  # noqa: BLE001
        ex["bam_realign"] = {"error": repr(e)}
    return ex


def pipeline_measurements():
    """loci/s of the unmodified reference pipeline on ssw.c vs under the wave scheduler (needs oracle/_ref_pipeline)"""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_pipeline as BP

        out = {}
        for cfg, loci in (("cfg1", 4), ("cfg3", 12)):
            out[cfg] = BP.measure(cfg, loci, workers=1, repeats=2)
        # zero-change use (SSW class swapped, no prefetch line, no wave scheduler): every miss is a device round trip
        out["cfg3_zero_change"] = BP.measure("cfg3", 4, workers=1, arms=("reference", "percall"))
        w = max(2, min(8, host_cores() // 2))
        # BASELINE configs[0] / [2] "in a BAM": the same loci written to BAM + BAI / FASTA + FAI and read back through the native
        # reader (indelpost_b200.bamio; pysam is absent) in BOTH arms
        out["cfg1_from_bam"] = BP.measure("cfg1", 4, workers=1, repeats=2, from_files=True)
        out["cfg3_from_bam"] = BP.measure("cfg3", 12, workers=1, repeats=2, from_files=True)
        # the same at equal host parallelism, through the product's locus-parallel driver (indelpost_b200.locuspool): ONE BAM +
        # FASTA with every locus, W worker processes in both arms (the wave arm's workers share GPU 0), started and warmed before
        # the timed map()
        out["cfg3_pool_%d_workers" % w] = BP.measure_pool("cfg3", 8 * w, workers=w)
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def ingest_measurements():
    """reads/s of pileup ingestion on one host core: the reference's make_pileup vs the native reader (tools/bench_ingest.py)"""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_ingest as BI

        return BI.measure(budget=0.7)
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import swbtest as T
    from indelpost_b200 import BatchAligner
    from indelpost_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    result_line = None
    if world > 1:
        torch.cuda.set_device(local)
        if not os.environ.get("SWB_BENCH_NO_NUMA_BIND"):
            bind_to_gpu_numa(local)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libswb200 has no CPU fallback)")
    dev = local if world > 1 else 0
    torch.cuda.set_device(dev)
    al = BatchAligner(dev)

    b = T.make_pairs_fast(args.pairs, READ_LEN, WIN_LEN, seed=1000 + rank)
    cells = b.cells()

    # pinned host staging for every input array (the batched entry point's contract): one-code-per-byte tables for the resident
    # leg and `e2e_codes`, 2-bit packed tables for the headline end-to-end leg
    keep, args_pos, _ = pinned_args(L, b, 0)
    keep2, args_pk, enc_pk = pinned_args(L, b, 2)
    kw = dict(mat=b.mat, n=5, score_size=2, flag=1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident leg --------------------------------------------------------------------------
    n = al.upload(*args_pos, **kw)
    for _ in range(args.warmup):
        al.compute()
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t0 = time.perf_counter()
    ev_ms = 0.0
    stage = {"ms_prepare": 0.0, "ms_forward": 0.0, "ms_reverse": 0.0, "ms_traceback": 0.0, "ms_band_round0": 0.0, "ms_band_rest": 0.0, "ms_certify": 0.0}
    launches = 0
    tm = {}
    for _ in range(args.steps):
        al.compute()
        tm = al.timing()
        ev_ms += tm["ms_total"]
        for k in stage:
            stage[k] += tm[k]
        launches += tm["n_launches"]
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    dt = t1 - t0
    res, arena = al.download(n)

    # ---- end-to-end legs (pinned host buffers -> results on host) ---------------------------------
    def e2e_leg(pos, enc):
        for _ in range(min(2, args.warmup)):
            al.align(*pos, copy=False, seq_encoding=enc, **kw)
        barrier()
        e0 = time.perf_counter()
        hh = dd = 0
        for _ in range(args.steps):
            al.align(*pos, copy=False, seq_encoding=enc, **kw)
            t = al.timing()
            hh, dd = t["h2d_bytes"], t["d2h_bytes"]
        barrier()
        return time.perf_counter() - e0, hh, dd

    dte_codes, h2d_codes, _ = e2e_leg(args_pos, L.SWB_SEQ_CODES)
    dte, h2d, d2h = e2e_leg(args_pk, enc_pk)

    if world > 1:
        tt = torch.tensor([dt, dte, dte_codes], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dte, dte_codes = float(tt[0]), float(tt[1]), float(tt[2])
        cc = torch.tensor([float(cells)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        total_cells = float(cc[0])
    else:
        total_cells = float(cells)

    if _ORIG_AFFINITY:
        os.sched_setaffinity(0, _ORIG_AFFINITY)           # the CPU legs below use every core the job was given
    if rank == 0:
        # sanity: results are real (spot-check against the CPU checker on a few pairs)
        sub = b.subset(np.arange(0, args.pairs, max(1, args.pairs // 64))[:64])
        ro, ao = (T.reference() if T.have_ref() else T.oracle()).align_batch(sub)
        idx = np.arange(0, args.pairs, max(1, args.pairs // 64))[:64]
        r_sub = res[idx].view(T.RESULT_DTYPE)
        for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "flag"):
            assert (r_sub[f] == ro[f]).all(), f"bench sanity check failed on {f}"

        value = total_cells * args.steps / dt / 1e9
        e2e = total_cells * args.steps / dte / 1e9
        pk = dpx_peak()
        peak_gcups = pk["lanes_per_clk_per_sm"] * pk["sms"] * pk["sm_mhz"] * 1e6 * 2 / pk["instr_per_cell_pair"] / 1e9
        # dominant kernel family = forward sweep; algorithmic cells per launch = sum(readLen*winLen) of this rank
        fwd_ms = stage["ms_forward"] / args.steps
        achieved = cells / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0
        alg_bytes = int(h2d_codes) + args.pairs * 40  # every input byte (one code per byte on the device) is read once by the sweep, one 40-byte result record written per pair
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            k = tj["kernels"].get("k_fast_fwd")
            if k:
                traffic = (k["dram_read_bytes"] + k["dram_write_bytes"]) / k["pairs"] * args.pairs
                traffic_src = tj["source"] + ": dram__bytes_read.sum + dram__bytes_write.sum of k_fast<19,0,false,0,8> (forward sweep), scaled per pair"
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback"
        cores = host_cores()
        sample_pairs = max(cores * 250, min(args.pairs, cores * 1500))      # the reference arm's sample
        sb = T.make_pairs_fast(sample_pairs, READ_LEN, WIN_LEN, seed=1234)
        cpu_align_parallel(sb.subset(np.arange(min(sample_pairs, cores * 100))), cores)      # warm-up
        cdt, kind = cpu_align_parallel(sb, cores)
        extra = None
        if world == 1 and not args.no_extra:
            extra = extra_measurements(al, L, T, peak_gcups)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (DPX) / u8+s16 striped emulation", "data": "synthetic", "config": workload_config(args.pairs, world),
            "reads_per_s": args.pairs * world * args.steps / dt,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "reads_per_s": args.pairs * world * args.steps / dte,
                    "ms_per_step": 1e3 * dte / args.steps, "host_buffers": "pinned, 2-bit packed sequence tables (SWB_SEQ_PACKED2)"},
            "e2e_codes": {"value": total_cells * args.steps / dte_codes / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_codes), "ms_per_step": 1e3 * dte_codes / args.steps,
                          "host_buffers": "pinned, one code per byte (the round-1 e2e figure)"},
            "gpu_launches": int(launches),
            "device_event_ms_per_step": ev_ms / args.steps,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "path_split": {"n_fast": int(tm.get("n_fast", 0)), "n_exact": int(tm.get("n_exact", 0))},
            "roofline": {"bound": "dpx", "kernel": "forward sweep (score/end/sub-optimal)", "achieved": achieved, "peak": peak_gcups, "unit": "GCUPS",
                         "frac": achieved / peak_gcups if peak_gcups else None, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                         "peak_source": pk.get("source", "profiles/dpx_peak.json"),
                         "hbm": {"achieved_gbs": alg_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else 0.0, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                                 "note": "algorithmic bytes = inputs read once (~0.009 B per cell): the path is integer-issue bound, not HBM bound"}},
            "cpu_baseline": {"value": sb.cells() / cdt / 1e9, "unit": UNIT, "cores": cores, "kind": kind, "reads_per_s": sample_pairs / cdt,
                             "sample": f"{sample_pairs} pairs of the same workload, split evenly over {cores} threads (one warm-up pass first); the reference arm uses the same sample"},
        }
        if extra is not None:
            line["extra"] = extra
        result_line = line
    if world == 1 and rank == 0 and not args.no_extra and result_line is not None:
        result_line.setdefault("extra", {})["pipeline"] = pipeline_measurements()
        result_line["extra"]["ingest"] = ingest_measurements()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result_line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the extra configs / mixes / pipeline block (1-GPU runs add it by default)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line (rank 0): while the benchmark runs, file descriptor 1 points at stderr, so that
    # banners printed by native libraries (NCCL's version line, the reference's warnings) cannot land in front of it
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
