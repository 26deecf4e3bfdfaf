#!/bin/bash
# Build variants of libitw_bcn.so with other unroll factors of the hottest loops (here, on the CPU box) and time them on
# the GPU:   bash tools/tune_unroll.sh build     then     gpurun -- 'bash tools/tune_unroll.sh time'
set -u
DIR=variants
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared -cudart static"
SRC=intel-texture-works-plugin_b200/csrc/itw_bcn.cu
# round 1 (BC7 slow / BC6H slow, 4096^2): assign 8 -> 36.24 ms, 16 -> 35.50, 4 -> 36.24; power 2 -> 36.24, 1 -> 36.83, 4 -> 36.08;
# BC6H assign 2 -> 14.42 ms, 1 -> 15.62, 4 -> 14.52, 8 -> 14.22.  Defaults: 16 / 4 / 8.  Round 2 re-run: profiles/r2_tune_unroll.txt.
VARIANTS="a8:-DITW_BC7_ASSIGN_UNROLL=8 a4:-DITW_BC7_ASSIGN_UNROLL=4 p2:-DITW_BC7_POWER_UNROLL=2 p8:-DITW_BC7_POWER_UNROLL=8 h4:-DITW_BC6_ASSIGN_UNROLL=4 h16:-DITW_BC6_ASSIGN_UNROLL=16"
if [ "${1:-}" = build ]; then
    mkdir -p $DIR
    for v in $VARIANTS; do
        name=${v%%:*}; def=${v#*:}
        nvcc $FLAGS $def $SRC -o $DIR/libitw_bcn_$name.so &
    done
    wait
    ls -la $DIR
else
    t() { ITW_BCN_LIB=$1 timeout 200 python bench.py --format $2 --profile $3 --no-cpu --no-extras --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1'.split('/')[-1], '$2', '$3', d['ms_per_step'])"; }
    t $PWD/intel-texture-works-plugin_b200/libitw_bcn.so BC7 slow
    t $PWD/intel-texture-works-plugin_b200/libitw_bcn.so BC7 basic
    t $PWD/intel-texture-works-plugin_b200/libitw_bcn.so BC6H bc6h_slow
    for n in a8 a4 p2 p8; do t $PWD/$DIR/libitw_bcn_$n.so BC7 slow; t $PWD/$DIR/libitw_bcn_$n.so BC7 basic; done
    for n in h4 h16; do t $PWD/$DIR/libitw_bcn_$n.so BC6H bc6h_slow; done
fi
