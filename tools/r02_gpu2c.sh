#!/bin/bash
# 2-GPU A/B of the fused gather's depth when rows come over NVLink: rows per tile, CTAs per SM, bulk-copy flavour
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
out=$O/r02_ab_gather_depth_g2.txt
: > $out
port=29800
for w in ${WORKLOADS:-papers100M}; do
  for v in "SPP_GATHER_TILE_ROWS=64" "SPP_GATHER_TILE_ROWS=128" "SPP_GATHER_TILE_ROWS=256" "SPP_GATHER_TILE_ROWS=256 SPP_GATHER_CTAS_PER_SM=3" \
           "SPP_GATHER_TILE_ROWS=128 SPP_GATHER_CTAS_PER_SM=4" "SPP_GATHER_TILE_ROWS=64 SPP_GATHER_CTAS_PER_SM=4" \
           "SPP_GATHER_BULK=1" "SPP_GATHER_BULK=1 SPP_BULK_CTAS_PER_SM=1" "SPP_GATHER_BULK=1 SPP_BULK_STAGES=8 SPP_BULK_TILE=2048"; do
    port=$((port+1))
    echo "== $w $v" >> $out
    env $v $TR --master-port $port bench.py --gpus 2 --workload $w --steps 300 --warmup 30 --device-only 2>> $out.err | tail -1 | cut -c1-120 >> $out
  done
done
cat $out
