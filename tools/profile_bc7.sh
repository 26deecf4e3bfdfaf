set -u
OUT=gpurun_out; LIB=intel-texture-works-plugin_b200/libitw_bcn.so; mkdir -p $OUT
ncu --clock-control none --set full --import-source on -k regex:bc7_kernel -c 1 -f -o $OUT/r1_final4_bc7_slow python bench.py --format BC7 --profile slow --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python profiles/summarise.py $OUT/r1_final4_bc7_slow.ncu-rep bc7_kernel $LIB bc7_kernelILb1 > $OUT/r1_final4_bc7_slow_ncu.txt
rm -f $OUT/r1_final4_bc7_slow.ncu-rep
timeout 200 python bench.py 2>&1 | tail -1 > $OUT/bench_default.json
python -c "import json; d=json.load(open('$OUT/bench_default.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
