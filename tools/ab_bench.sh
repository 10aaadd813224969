#!/bin/bash
# A/B runs of bench.py with LH2B_SET_<setting> overrides: tools/ab_bench.sh "widePrefetch=1" "widePrefetch=2 shadeBlocks=6" ...
mkdir -p gpurun_out
run() {
  local tag="$1"; shift
  env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('%-40s value %8.1f e2e %8.1f ms %.4f | ge %.4f shade %.4f conn %.4f' % ('$tag', d['value'], d['e2e']['value'], d['ms_per_step'], s['generateExtendMs'], s['shadeMs'], s['connectMs']))"
}
run base
for cfg in "$@"; do
  args=(); for kv in $cfg; do args+=("LH2B_SET_$kv"); done
  run "$cfg" "${args[@]}"
done
run base-again
