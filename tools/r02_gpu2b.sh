#!/bin/bash
# 2-GPU A/B of the class-split gather (peer buckets on a side stream) against the fused gather
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_tests_g2b.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_tests_g2b.log
: > $O/r02_ab_gather_split_g2.txt
port=29700
for w in papers100M products; do
  for v in "SPP_GATHER_SPLIT=1" "SPP_GATHER_SPLIT=0" "SPP_GATHER_SPLIT=1 SPP_GATHER_BULK=1"; do
    port=$((port+1))
    echo "== $w $v" >> $O/r02_ab_gather_split_g2.txt
    env $v $TR --master-port $port bench.py --gpus 2 --workload $w --steps 300 --warmup 30 --device-only 2>> $O/r02_ab_gather_split_g2.err | tail -1 | cut -c1-160 >> $O/r02_ab_gather_split_g2.txt
  done
done
cat $O/r02_ab_gather_split_g2.txt
