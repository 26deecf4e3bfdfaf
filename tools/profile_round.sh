#!/bin/bash
# ncu captures behind the summaries in profiles/ (run on the GPU box through gpurun; results land in gpurun_out/).
#   bash tools/profile_round.sh <tag>            e.g. r1_final
# 1. launch list of the default bench command (per-launch gpu__time_duration, --clock-control none)
# 2. one --set full capture of the dominant kernel of BC7 slow, BC6H slow, BC1, BC3 at 4096^2
# 3. --set full captures of the decode and front-end kernels
# Each report is summarised ON THE BOX by profiles/summarise.py (gpurun brings back at most 64 MiB); only the BC7
# report itself is kept.  Numbers printed by bench.py under ncu are never bench values.
set -u
TAG=${1:-r1_final}
OUT=gpurun_out
LIB=intel-texture-works-plugin_b200/libitw_bcn.so
mkdir -p $OUT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_default_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${TAG}_launches_bench_stdout.log 2>&1
full() {  # name, kernel regex, count, import-source(0/1), section, command...
    local name=$1 kern=$2 count=$3 src=$4 section=$5; shift 5
    local extra=""
    [ "$src" = 1 ] && extra="--import-source on"
    $NCU --set full $extra -k regex:$kern -c $count -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
    if [ "$src" = 1 ]; then python profiles/summarise.py $OUT/${TAG}_$name.ncu-rep $kern $LIB $section > $OUT/${TAG}_${name}_ncu.txt 2>> $OUT/${TAG}_$name.log
    else python profiles/summarise.py $OUT/${TAG}_$name.ncu-rep $kern > $OUT/${TAG}_${name}_ncu.txt 2>> $OUT/${TAG}_$name.log; fi
    [ "$name" = bc7_slow ] || rm -f $OUT/${TAG}_$name.ncu-rep
}
full bc7_slow bc7_kernel 1 1 bc7_kernelILb1 python bench.py --format BC7 --profile slow --steps 1 --warmup 1 --no-cpu
full bc6h_slow bc6h_kernel 1 1 bc6h_kernelILb1 python bench.py --format BC6H --profile bc6h_slow --steps 1 --warmup 1 --no-cpu
full bc1 bc1_bc3_kernel 1 1 bc1_bc3_kernelILb0ELb1 python bench.py --format BC1 --steps 1 --warmup 1 --no-cpu
full bc3 bc1_bc3_kernel 1 1 bc1_bc3_kernelILb1ELb1 python bench.py --format BC3 --steps 1 --warmup 1 --no-cpu
full decode decode_kernel 12 0 - python tools/decode_bench.py --reps 1 --warm 1
full front front_kernel 14 0 - python tools/frontend_bench.py --reps 1 --warm 1
ls -la $OUT | tail -24
