for v in "traversalVariant=0" "triThreshold=1" "triThreshold=4" "triThreshold=8" "triThreshold=12" "triThreshold=16" "triThreshold=24" "triThreshold=32" "triThresholdShadow=1" "triThresholdShadow=6" "triThresholdShadow=20" "refillThreshold=1" "refillThreshold=4" "refillThreshold=16" "refillThreshold=32" "wideBlocksPerSM=4" "wideBlocksPerSM=6" "wideBlocksPerSM=16"; do
  k=${v%%=*}; val=${v##*=}
  echo "== $v"; env LH2B_SET_$k=$val python tools/trace_probe.py 2>&1 | grep -E "x10"
done
