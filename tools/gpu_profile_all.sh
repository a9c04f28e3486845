#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): `ncu --set full` captures that cover every kernel of the path, condensed on the box
# (tools/ncu_summarize.py) because the reports themselves exceed gpurun's return limit; only the report of the dominant
# kernel (forward sweep, 2 launches) is brought back whole for the source page.
# usage: tools/gpu_profile_all.sh <tag>      outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-r02}
OUT=gpurun_out
TMP=/tmp/swb_ncu
mkdir -p $OUT $TMP
NCU="ncu --clock-control none"
cap() {   # cap <name> <launch-count> <command...>
    local name=$1 cnt=$2; shift 2
    $NCU --set full --import-source on -c $cnt -o $TMP/${TAG}_$name -f "$@" > $OUT/${TAG}_ncu_$name.log 2>&1
    ncu -i $TMP/${TAG}_$name.ncu-rep --page raw --csv > $TMP/${TAG}_${name}_raw.csv 2>> $OUT/${TAG}_ncu_$name.log
    python tools/ncu_summarize.py $TMP/${TAG}_${name}_raw.csv > $OUT/${TAG}_${name}_summary.json 2>> $OUT/${TAG}_ncu_$name.log
}
# 1. headline workload (cfg2), 200 k pairs: forward sweep, banded reverse, register-band traceback, certificate
python bench.py --no-extra --pairs 200000 --steps 1 --warmup 1 > $OUT/${TAG}_plain_cfg2.json 2> $OUT/${TAG}_plain_cfg2.err || exit 1
cap cfg2 120 python bench.py --no-extra --pairs 200000 --steps 1 --warmup 0
# 2. indelPost penalty mix (ge = 0, go = len(read)): wavefront reverse, wide bands (k_band_warp, k_band), exact kernels
python tools/bench_grid_mix.py 200000 > $OUT/${TAG}_plain_mix.json 2> $OUT/${TAG}_plain_mix.err || exit 1
cap mix 200 python tools/profile_mix_once.py 200000
# 3. short reads (unsafe zone of the 8-bit pass) + indel extraction
python tools/profile_short_once.py 60000 > $OUT/${TAG}_plain_short.json 2> $OUT/${TAG}_plain_short.err || exit 1
cap short 120 python tools/profile_short_once.py 60000
# the dominant kernel alone, whole report (source page, stall reasons)
$NCU --set full --import-source on -k regex:k_fast -c 2 -o $OUT/${TAG}_kfast -f python bench.py --no-extra --pairs 200000 --steps 1 --warmup 0 > $OUT/${TAG}_ncu_kfast.log 2>&1
ls -la $OUT/${TAG}_* $TMP
