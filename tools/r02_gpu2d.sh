#!/bin/bash
# 2-GPU check of the gather CTA count chosen at 4 GPUs
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
out=$O/r02_ab_nvlink_regime_g2.txt
: > $out
port=30200
for v in "SPP_X=0" "SPP_GATHER_CTAS_PER_SM=1" "SPP_X=1" "SPP_GATHER_CTAS_PER_SM=1 SPP_GATHER_TILE_ROWS=256"; do
  port=$((port+1))
  echo "== $v" >> $out
  env $v $TR --master-port $port bench.py --gpus 2 --steps 300 --warmup 30 --device-only 2>> $out.err | tail -1 | cut -c1-110 >> $out
done
cat $out
