"""Condense an `ncu --page raw --csv` export into one record per kernel (demangled base name + template arguments):
launches, total / mean duration, registers, shared memory, achieved occupancy, issue rate, ALU / FMA / LSU pipe utilisation and
DRAM bytes per launch.  Runs on the GPU box right after the capture, so only the small JSON travels back (a `--set full` report of
a hundred launches is larger than gpurun's 64 MiB return limit).

    ncu -i x.ncu-rep --page raw --csv > x_raw.csv ; python tools/ncu_summarize.py x_raw.csv > x_summary.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "launch__registers_per_thread": "regs",
    "launch__block_size": "block",
    "launch__grid_size": "grid",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "sm__inst_issued.avg.per_cycle_active": "issued_per_cycle_active",
    "sm__inst_executed.sum": "inst_executed",
    "sm__inst_executed_pipe_alu.sum": "inst_alu",
    "sm__inst_executed_pipe_fma.sum": "inst_fma",
    "sm__inst_executed_pipe_lsu.sum": "inst_lsu",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_inst_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_inst_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_inst_pct",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard_pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct": "stall_short_scoreboard_pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct": "stall_math_pipe_pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct": "stall_wait_pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct": "stall_not_selected_pct",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    # the header is the first row that contains "Kernel Name"; the next row holds the units
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units = rows[h], rows[h + 1]
    col = {name: i for i, name in enumerate(head)}
    kn = col["Kernel Name"]
    out = {}
    for r in rows[h + 2:]:
        if len(r) <= kn:
            continue
        name = re.sub(r"\(.*$", "", r[kn]).strip()
        rec = out.setdefault(name, {"launches": 0})
        rec["launches"] += 1
        for metric, key in WANT.items():
            if metric in col:
                v = num(r[col[metric]])
                if v is None:
                    continue
                u = units[col[metric]]
                if key == "duration_ns":
                    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
                if key in ("dram_read_bytes", "dram_write_bytes", "l2_bytes", "smem_dyn", "smem_static"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                rec.setdefault("_" + key, []).append(v)
    res = {}
    for name, rec in out.items():
        o = {"launches": rec["launches"]}
        for k, vs in rec.items():
            if not k.startswith("_"):
                continue
            key = k[1:]
            if key in ("duration_ns", "inst_executed", "inst_alu", "inst_fma", "inst_lsu", "dram_read_bytes", "dram_write_bytes", "l2_bytes"):
                o[key + "_total"] = sum(vs)
                o[key + "_mean"] = sum(vs) / len(vs)
            elif key in ("regs", "block", "smem_dyn", "smem_static"):
                o[key] = max(vs)
            elif key == "grid":
                o["grid_max"] = max(vs)
            else:
                # duration-weighted mean where durations are known, else plain mean
                d = rec.get("_duration_ns")
                if d and len(d) == len(vs) and sum(d) > 0:
                    o[key] = sum(a * b for a, b in zip(vs, d)) / sum(d)
                else:
                    o[key] = sum(vs) / len(vs)
        res[name] = o
    tot = sum(o.get("duration_ns_total", 0) for o in res.values())
    for o in res.values():
        if tot:
            o["share_of_captured_time"] = o.get("duration_ns_total", 0) / tot
    json.dump({"source": sys.argv[1], "kernels": dict(sorted(res.items(), key=lambda kv: -kv[1].get("duration_ns_total", 0)))}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
