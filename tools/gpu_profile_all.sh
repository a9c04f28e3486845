#!/bin/bash
# Runs ON THE GPU BOX (through gpurun): ncu launch lists + `--set full` captures that cover every kernel of the path.
# usage: tools/gpu_profile_all.sh <tag>      outputs under gpurun_out/<tag>_*
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# 1. headline workload (cfg2), 200 k pairs: forward sweep, banded reverse, register-band traceback, certificate
python bench.py --pairs 200000 --steps 1 --warmup 1 > $OUT/${TAG}_plain_cfg2.json 2> $OUT/${TAG}_plain_cfg2.err || exit 1
$NCU --set full --import-source on -c 120 -o $OUT/${TAG}_cfg2_full -f python bench.py --pairs 200000 --steps 1 --warmup 0 > $OUT/${TAG}_ncu_cfg2.log 2>&1
# 2. indelPost penalty mix (ge = 0, go = len(read)): wavefront reverse, wide bands (k_band_warp, k_band), exact kernels
python tools/bench_grid_mix.py 200000 > $OUT/${TAG}_plain_mix.json 2> $OUT/${TAG}_plain_mix.err || exit 1
$NCU --set full --import-source on -c 200 -o $OUT/${TAG}_mix_full -f python tools/profile_mix_once.py 200000 > $OUT/${TAG}_ncu_mix.log 2>&1
# 3. short reads (unsafe zone of the 8-bit pass) + indel extraction
python tools/profile_short_once.py 60000 > $OUT/${TAG}_plain_short.json 2> $OUT/${TAG}_plain_short.err || exit 1
$NCU --set full --import-source on -c 120 -o $OUT/${TAG}_short_full -f python tools/profile_short_once.py 60000 > $OUT/${TAG}_ncu_short.log 2>&1
ls -la $OUT/${TAG}_*
