#!/bin/bash
# A/B runs of the device-timed pipeline under different SPP_* switches (one process per variant).
# usage: tools/ab.sh out_file "ENV1=.. ENV2=..|--bench-args" "ENV..|" ...
out=$1; shift
: > "$out"
for v in "$@"; do
  envs="${v%%|*}"; args=""
  [[ "$v" == *"|"* ]] && args="${v#*|}"
  echo "== $v" >> "$out"
  env $envs python bench.py --device-only --steps ${AB_STEPS:-300} --warmup 30 ${AB_ARGS:-} $args 2>> "$out.err" | sed -e 's/"env".*//' >> "$out"
done
cat "$out"
