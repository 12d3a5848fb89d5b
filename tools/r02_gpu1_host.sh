#!/bin/bash
# Round-2 session: native host path (csrc/host_session.cpp) -- GPU suite, then A/B of the public-API rate
# with the interpreter path (SPP_NATIVE_HOST=0) on the host-bound shapes.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
(time timeout 240 python -m pytest tests -m gpu -x -q) > $O/r02_tests_host_g1.log 2>&1
echo "pytest rc=$?" >> $O/r02_tests_host_g1.log
tail -6 $O/r02_tests_host_g1.log
for w in arxiv products-layerwise; do
  for nat in 1; do
    SPP_NATIVE_HOST=$nat timeout 120 python bench.py --workload $w --steps 400 --warmup 20 --no-cpu-baseline \
      > $O/r02_host_${w}_native${nat}.json 2> $O/r02_host_${w}_native${nat}.err
    echo "== $w native=$nat rc=$?"; python - <<P
import json
try:
    d = json.loads(open("$O/r02_host_${w}_native${nat}.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["per_batch_us"], d["parity"]["ok"])
except Exception as e:
    print("no line:", e)
P
    tail -2 $O/r02_host_${w}_native${nat}.err
  done
done
timeout 120 python bench.py --workload arxiv --steps 400 --warmup 20 --no-cpu-baseline --profile-e2e \
  > $O/r02_host_arxiv_profile.json 2> $O/r02_host_arxiv_profile.txt
head -30 $O/r02_host_arxiv_profile.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_host_papers_g1_s20.json 2> $O/r02_host_papers_g1_s20.err
echo "papers rc=$?"; head -c 700 $O/r02_host_papers_g1_s20.json; tail -2 $O/r02_host_papers_g1_s20.err
