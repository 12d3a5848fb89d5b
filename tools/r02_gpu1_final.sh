#!/bin/bash
# Round-2 closing single-GPU session: GPU suite + smoke with the native host path, then the bench lines
# of the host-bound shapes (arxiv, layer-wise inference) with their CPU baselines, a deeper-pipeline A/B
# of the layer-wise shape, and the default (papers100M-shaped) line.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 240 python -m pytest tests -m gpu -x -q) > $O/r02_tests_final_g1.log 2>&1
echo "pytest rc=$?" >> $O/r02_tests_final_g1.log
tail -6 $O/r02_tests_final_g1.log
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke_final.log
show() { python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline") or {}
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["per_batch_us"].get("p50"),
          "cpu", cb.get("value"), "parity", d["parity"]["ok"], "roofline", d["roofline"]["frac"])
except Exception as e:
    print("no line:", e)
P
}
timeout 200 python bench.py --workload products-layerwise --steps 400 --warmup 20 > $O/r02_bench_products_layerwise_g1.json 2> $O/r02_bench_products_layerwise_g1.err
echo "== layerwise rc=$?"; show $O/r02_bench_products_layerwise_g1.json
timeout 200 python bench.py --workload arxiv --steps 400 --warmup 20 > $O/r02_bench_arxiv_g1.json 2> $O/r02_bench_arxiv_g1.err
echo "== arxiv rc=$?"; show $O/r02_bench_arxiv_g1.json
for d in 10; do
  SPP_SESSION_DEPTH=$d timeout 100 python bench.py --workload products-layerwise --depth $d --steps 400 --warmup 20 --no-cpu-baseline \
    > $O/r02_host_layerwise_depth$d.json 2> $O/r02_host_layerwise_depth$d.err
  echo "== layerwise depth $d rc=$?"; show $O/r02_host_layerwise_depth$d.json
done
timeout 300 python bench.py --steps 200 --warmup 20 > $O/r02_bench_papers_g1.json 2> $O/r02_bench_papers_g1.err
echo "== papers rc=$?"; show $O/r02_bench_papers_g1.json
