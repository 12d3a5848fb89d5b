#!/bin/bash
# usage: tools/gather_ab.sh out_file "ENV..|--shape papers" ...
out=$1; shift
: > "$out"
for v in "$@"; do
  envs="${v%%|*}"; args=""
  [[ "$v" == *"|"* ]] && args="${v#*|}"
  env $envs python tools/gather_bench.py $args 2>> "$out.err" >> "$out"
done
cat "$out"
