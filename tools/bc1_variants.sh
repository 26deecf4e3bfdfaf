#!/bin/bash
# BC1 / BC3 register budget variants: bash tools/bc1_variants.sh build   (here), then   gpurun -- 'bash tools/bc1_variants.sh'
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared -cudart static"
if [ "${1:-}" = build ]; then
    mkdir -p variants
    for m in 3 4 5 6; do nvcc $FLAGS "-DITW_BC1_MIN_CTAS(alpha)=$m" intel-texture-works-plugin_b200/csrc/itw_bcn.cu -o variants/libitw_bcn_m$m.so & done
    wait; exit 0
fi
t() { ITW_BCN_LIB=$1 timeout 200 python bench.py --format $2 --no-cpu --no-extras --steps 10 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1'.split('/')[-1], '$2', d['ms_per_step'])"; }
for rep in 1 2; do
for f in BC1 BC3; do
  t $PWD/intel-texture-works-plugin_b200/libitw_bcn.so $f
  for m in 3 4 5 6; do [ -f variants/libitw_bcn_m$m.so ] && t $PWD/variants/libitw_bcn_m$m.so $f; done
done; done
