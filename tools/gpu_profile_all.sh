#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): ncu captures that cover every kernel of the path, condensed on the box
# (tools/ncu_summarize.py) because the reports themselves exceed gpurun's return limit.  A fixed metric list (a few passes per
# launch) for the survey of all kernels; `--set full` only for two launches of the dominant kernel (source page, stall reasons).
# usage: tools/gpu_profile_all.sh <tag> [cfg2|mix|short|kfast ...]     outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-r02}; shift || true
WHAT=${*:-cfg2 mix short kfast}
OUT=gpurun_out
TMP=/tmp/swb_ncu
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
METRICS=gpu__time_duration.sum,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active,sm__inst_issued.avg.per_cycle_active,sm__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct
cap() {   # cap <name> <launch-count> <command...>
    local name=$1 cnt=$2; shift 2
    $NCU --metrics $METRICS -c $cnt -o $TMP/${TAG}_$name -f "$@" > $OUT/${TAG}_ncu_$name.log 2>&1
    ncu -i $TMP/${TAG}_$name.ncu-rep --page raw --csv > $TMP/${TAG}_${name}_raw.csv 2>> $OUT/${TAG}_ncu_$name.log
    python tools/ncu_summarize.py $TMP/${TAG}_${name}_raw.csv > $OUT/${TAG}_${name}_summary.json 2>> $OUT/${TAG}_ncu_$name.log
    grep -c "Profiling" $OUT/${TAG}_ncu_$name.log
}
for w in $WHAT; do
  case $w in
    cfg2)   # headline workload: forward sweep, banded reverse, register-band traceback, certificate, sandwich verification
      cap cfg2 160 python bench.py --no-extra --pairs 200000 --steps 1 --warmup 0 ;;
    mix)    # indelPost penalty mix (ge = 0, go = len(read)): wavefront reverse, wide bands (k_band_warp, k_band), exact kernels
      cap mix 220 python tools/profile_mix_once.py 200000 ;;
    short)  # short reads (sandwich sweep, exact kernels for the rest) + indel extraction
      cap short 160 python tools/profile_short_once.py 100000 ;;
    kfast)  # the dominant kernel alone, whole report
      $NCU --set full --import-source on -k regex:k_fast -c 2 -o $OUT/${TAG}_kfast -f python bench.py --no-extra --pairs 200000 --steps 1 --warmup 0 > $OUT/${TAG}_ncu_kfast.log 2>&1 ;;
  esac
done
ls -la $OUT/${TAG}_* $TMP | head -40
