#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun; results land in gpurun_out/, summaries are copied to profiles/).
#   bash tools/profile_r2.sh <tag> [bc7|bc6h|bc1|bc3|launches ...]
set -u
TAG=${1:-r2}; shift
WHAT=${*:-bc7 bc6h bc1 bc3 bc45 mips decode front launches}
OUT=gpurun_out
LIB=intel-texture-works-plugin_b200/libitw_bcn.so
mkdir -p $OUT
NCU="ncu --clock-control none"
full() {  # name, kernel regex, mangled-section, command...
    local name=$1 kern=$2 section=$3; shift 3
    $NCU --set full --import-source on -k regex:$kern -c 1 -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
    python profiles/summarise.py $OUT/${TAG}_$name.ncu-rep $kern $LIB $section > $OUT/${TAG}_${name}_ncu.txt 2>> $OUT/${TAG}_$name.log
    [ "$name" = bc7_slow ] || [ "$name" = bc6h_slow ] || rm -f $OUT/${TAG}_$name.ncu-rep
}
many() {  # name, kernel regex, count, command...: --set full without source, every matching launch summarised
    local name=$1 kern=$2 count=$3; shift 3
    $NCU --set full -k regex:"$kern" -c $count -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
    python profiles/summarise.py $OUT/${TAG}_$name.ncu-rep "" > $OUT/${TAG}_${name}_ncu.txt 2>> $OUT/${TAG}_$name.log
    rm -f $OUT/${TAG}_$name.ncu-rep
}
for w in $WHAT; do
  case $w in
    launches) $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/${TAG}_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > $OUT/${TAG}_launches_bench_stdout.log 2>&1 ;;
    bc7)  full bc7_slow bc7_kernel bc7_kernelILb1 python bench.py --format BC7 --profile slow --steps 1 --warmup 1 --no-cpu --no-extras ;;
    bc6h) full bc6h_slow bc6h_kernel bc6h_kernelILb1 python bench.py --format BC6H --profile bc6h_slow --steps 1 --warmup 1 --no-cpu --no-extras ;;
    bc1)  full bc1 bc1_bc3_kernel bc1_bc3_kernelILb0ELb1 python bench.py --format BC1 --steps 1 --warmup 1 --no-cpu --no-extras ;;
    bc3)  full bc3 bc1_bc3_kernel bc1_bc3_kernelILb1ELb1 python bench.py --format BC3 --steps 1 --warmup 1 --no-cpu --no-extras ;;
    bc45) many bc4 bc4_bc5_kernel 1 python bench.py --format BC4 --steps 1 --warmup 1 --no-cpu --no-extras
          many bc5 bc4_bc5_kernel 1 python bench.py --format BC5 --steps 1 --warmup 1 --no-cpu --no-extras ;;
    mips) many mips "mip_|pad_" 8 python tools/mip_chain.py --reps 1 ;;
    decode) many decode decode_kernel 12 python tools/decode_bench.py --reps 1 --warm 1 ;;
    front) many front front_kernel 14 python tools/frontend_bench.py --reps 1 --warm 1 ;;
  esac
done
ls -la $OUT | tail -12
