#!/bin/bash
# Device-resident ms per 4096x4096 surface and end-to-end Mtexels/s of every encoder configuration worth quoting, one line each.
#   bash tools/time_profiles.sh > gpurun_out/<tag>_timings.txt
run() {  # format [profile]
    local args="--format $1"; [ -n "${2:-}" ] && args="$args --profile $2"
    python bench.py $args --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', '${2:-}', 'ms', round(d['ms_per_step'], 4), 'Mtexels/s', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'sm_mhz', d['clocks'].get('sm_mhz'))"
}
for p in slow basic fast veryfast ultrafast alpha_slow alpha_basic alpha_fast alpha_veryfast alpha_ultrafast; do run BC7 $p; done
for p in bc6h_veryslow bc6h_slow bc6h_basic bc6h_fast bc6h_veryfast; do run BC6H $p; done
run BC1; run BC3; run BC4; run BC5
