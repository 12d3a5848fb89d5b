#!/bin/bash
# multi-GPU A/B of the device-timed pipeline: tools/ab_multi.sh NGPUS out_file "ENV..|--bench-args" ...
n=$1; out=$2; shift 2
: > "$out"
port=29600
for v in "$@"; do
  envs="${v%%|*}"; args=""
  [[ "$v" == *"|"* ]] && args="${v#*|}"
  port=$((port+1))
  echo "== $v" >> "$out"
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --device-only --steps ${AB_STEPS:-300} --warmup 30 $args 2>> "$out.err" | grep device_only | sed -e 's/"env".*//' >> "$out"
done
cat "$out"
