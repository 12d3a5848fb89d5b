#!/bin/bash
# closing check: GPU suite (incl. native vs interpreter host path) + the layer-wise line with consumer-blocked statistics
set -u
O=gpurun_out
mkdir -p $O
(time timeout 240 python -m pytest tests -m gpu -x -q) > $O/r02_tests_check_g1.log 2>&1
echo "pytest rc=$?" >> $O/r02_tests_check_g1.log
tail -6 $O/r02_tests_check_g1.log
timeout 100 python bench.py --workload products-layerwise --steps 400 --warmup 20 --no-cpu-baseline > $O/r02_check_layerwise.json 2> $O/r02_check_layerwise.err
echo "layerwise rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_check_layerwise.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["per_batch_us"], d["parity"]["ok"])
P
