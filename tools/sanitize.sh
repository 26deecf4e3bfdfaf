#!/bin/bash
# compute-sanitizer over __graft_entry__.smoke() (every encoder incl. the TMA-staged kernels, front end, decoder) and over a
# multi-device / deferred-mode exercise; logs land in gpurun_out/ and are copied to profiles/.
#   gpurun -- 'bash tools/sanitize.sh r2'
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > $OUT/${TAG}_sanitizer_${tool}_smoke.log 2>&1
    echo "$tool smoke: exit $?" >> $OUT/${TAG}_sanitizer_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|hazard" $OUT/${TAG}_sanitizer_${tool}_smoke.log | tail -3 >> $OUT/${TAG}_sanitizer_summary.txt
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_multi_device.py -m gpu -q -p no:cacheprovider -k "zero_height or wrappers or deferred_mode or fanout" > $OUT/${TAG}_sanitizer_memcheck_host_paths.log 2>&1
echo "memcheck host paths: exit $?" >> $OUT/${TAG}_sanitizer_summary.txt
grep -E "ERROR SUMMARY|passed|failed" $OUT/${TAG}_sanitizer_memcheck_host_paths.log | tail -3 >> $OUT/${TAG}_sanitizer_summary.txt
# BC7 in rounds of sixteen blocks per warp (surfaces of >= 37 888 blocks: aliased chain-phase storage, single TMA tile buffer)
for tool in memcheck racecheck; do
    timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "sixteen_block" > $OUT/${TAG}_sanitizer_${tool}_bc7_rounds16.log 2>&1
    echo "$tool BC7 sixteen-block rounds: exit $?" >> $OUT/${TAG}_sanitizer_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" $OUT/${TAG}_sanitizer_${tool}_bc7_rounds16.log | tail -3 >> $OUT/${TAG}_sanitizer_summary.txt
done
cat $OUT/${TAG}_sanitizer_summary.txt
