#!/bin/bash
# 4-GPU A/B of the NVLink-bound regime: gather CTAs per SM / tile height, batches in flight
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
out=$O/r02_ab_nvlink_regime_g4.txt
: > $out
port=30100
run() {  # env... | args
  port=$((port+1))
  echo "== $1 | $2" >> $out
  env $1 $TR --master-port $port bench.py --gpus 4 --steps 300 --warmup 30 --device-only $2 2>> $out.err | tail -1 | cut -c1-110 >> $out
}
run "SPP_X=0" ""
run "SPP_GATHER_CTAS_PER_SM=1 SPP_GATHER_TILE_ROWS=256" ""
run "SPP_GATHER_CTAS_PER_SM=1 SPP_GATHER_TILE_ROWS=128" ""
run "SPP_GATHER_TILE_ROWS=64" ""
run "SPP_X=0" "--depth 4"
run "SPP_X=0" "--depth 8"
run "SPP_GATHER_CTAS_PER_SM=3 SPP_GATHER_TILE_ROWS=64" ""
cat $out
