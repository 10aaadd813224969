# C5 (4K filtered frames, tile-sharded over N GPUs): tail on rank 0 vs the sharded filter chain. usage: bash tools/c5_modes.sh N [modes...]
N=$1; shift
MODES=${@:-root shard_il shard_ct}
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/run_c5_scaling.py gpurun_out/c5_n${N}_$tag.json 40 device 2>&1 | grep '^{\|tile timing' | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$tag', 'N=$N', round(d['ms_per_frame'], 3), 'ms', d['image_mean'])
    else:
        print('  ', l.strip())"; }
for m in $MODES; do
  case $m in
    root) run root LH2B_SET_tileFilterShard=0 ;;
    shard_il) run shard_il LH2B_SET_tileFilterShard=1 LH2B_SET_tileInterleave=1 ;;
    shard_ct) run shard_ct LH2B_SET_tileFilterShard=1 LH2B_SET_tileInterleave=0 ;;
    shard_il_t) run shard_il_t LH2B_TILE_TIMING=1 LH2B_SET_tileFilterShard=1 LH2B_SET_tileInterleave=1 ;;
  esac
done
