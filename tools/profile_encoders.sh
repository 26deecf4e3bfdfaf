set -u
TAG=r1_final3
OUT=gpurun_out
LIB=intel-texture-works-plugin_b200/libitw_bcn.so
mkdir -p $OUT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
$NCU --set full --import-source on -k regex:bc7_kernel -c 1 -f -o $OUT/${TAG}_bc7_slow python bench.py --format BC7 --profile slow --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python profiles/summarise.py $OUT/${TAG}_bc7_slow.ncu-rep bc7_kernel $LIB bc7_kernelILb1 > $OUT/${TAG}_bc7_slow_ncu.txt
rm -f $OUT/${TAG}_bc7_slow.ncu-rep
$NCU --set full --import-source on -k regex:bc6h_kernel -c 1 -f -o $OUT/${TAG}_bc6h_slow python bench.py --format BC6H --profile bc6h_slow --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
python profiles/summarise.py $OUT/${TAG}_bc6h_slow.ncu-rep bc6h_kernel $LIB bc6h_kernelILb1 > $OUT/${TAG}_bc6h_slow_ncu.txt
rm -f $OUT/${TAG}_bc6h_slow.ncu-rep
ls $OUT | grep final3
