#!/bin/bash
# Round-2 closing two-GPU session (gpurun --gpus 2): the P2P path through the native host path.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time timeout 120 python -m pytest tests/test_gpu_multi.py -x -q) > $O/r02_tests_final_g2.log 2>&1
echo "pytest rc=$?" >> $O/r02_tests_final_g2.log; tail -4 $O/r02_tests_final_g2.log
timeout 200 $TR --master-port 29711 bench.py --gpus 2 --steps 200 --warmup 20 > $O/r02_bench_papers_g2.json 2> $O/r02_bench_papers_g2.err
echo "papers g2 rc=$?"; head -c 300 $O/r02_bench_papers_g2.json; echo; tail -2 $O/r02_bench_papers_g2.err
timeout 120 $TR --master-port 29712 bench.py --gpus 2 --workload products-layerwise --steps 400 --warmup 20 > $O/r02_bench_products_layerwise_g2.json 2> $O/r02_bench_products_layerwise_g2.err
echo "layerwise g2 rc=$?"; head -c 300 $O/r02_bench_products_layerwise_g2.json; echo; tail -2 $O/r02_bench_products_layerwise_g2.err
python - <<'P'
import json
for f in ("r02_bench_papers_g2", "r02_bench_products_layerwise_g2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["parity"], d.get("nvlink_roofline", {}).get("frac"), d.get("rows_served"))
    except Exception as e:
        print(f, "no line", e)
P
