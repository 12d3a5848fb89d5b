#!/bin/bash
# Round-2 single-GPU measurement session (run on the B200 box through gpurun); everything lands in gpurun_out/.
#   tools/r02_gpu1.sh [tests] [bench] [ab] [ncu]
set -u
O=gpurun_out
mkdir -p $O
what="${*:-tests bench ab ncu}"
if [[ $what == *tests* ]]; then
  (time python -m pytest tests -m gpu -x -q) > $O/r02_tests_g1.log 2>&1
  echo "pytest rc=$?" >> $O/r02_tests_g1.log
  tail -4 $O/r02_tests_g1.log
fi
if [[ $what == *bench* ]]; then
  python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_ref_papers_g1.json 2> $O/r02_bench_ref_papers_g1.err
  python bench.py --steps 200 --warmup 20 > $O/r02_bench_papers_g1.json 2> $O/r02_bench_papers_g1.err
  python bench.py --steps 20 --warmup 5 > $O/r02_bench_papers_g1_s20.json 2>> $O/r02_bench_papers_g1.err
  python bench.py --steps 200 --warmup 20 --emulate-peers --no-cpu-baseline > $O/r02_bench_papers_emulated_peers_g1.json 2> $O/r02_bench_papers_emulated_peers_g1.err
  python bench.py --workload products --steps 200 --warmup 20 > $O/r02_bench_products_g1.json 2> $O/r02_bench_products_g1.err
  python bench.py --workload products-layerwise --steps 200 --warmup 20 > $O/r02_bench_products_layerwise_g1.json 2> $O/r02_bench_products_layerwise_g1.err
  python bench.py --workload arxiv --steps 200 --warmup 20 > $O/r02_bench_arxiv_g1.json 2> $O/r02_bench_arxiv_g1.err
  for f in papers_g1 papers_g1_s20 papers_emulated_peers_g1 products_g1 products_layerwise_g1 arxiv_g1 ref_papers_g1; do
    echo "== $f"; head -c 600 $O/r02_bench_$f.json; echo; tail -3 $O/r02_bench_$f.err 2>/dev/null
  done
fi
if [[ $what == *ab* ]]; then
  python tools/gather_ab.py > $O/r02_ab_gather_bulk_g1.txt 2> $O/r02_ab_gather_bulk_g1.err
  cat $O/r02_ab_gather_bulk_g1.txt; tail -3 $O/r02_ab_gather_bulk_g1.err
fi
if [[ $what == *ncu* ]]; then
  B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-peers"
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k regex:k_ -c 400 --csv --log-file $O/r02_launches_warm_papers_emulated.csv $B > $O/ncu1.log 2>&1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_gather -s 21 -c 2 \
      -o $O/r02_gather_partitioned -f $B > $O/ncu2.log 2>&1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_split -s 30 -c 3 \
      -o $O/r02_split -f $B > $O/ncu3.log 2>&1
  B2="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k regex:k_ -c 400 --csv --log-file $O/r02_launches_warm_papers.csv $B2 > $O/ncu4.log 2>&1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_gather -s 20 -c 2 \
      -o $O/r02_gather_papers_local -f $B2 > $O/ncu5.log 2>&1
  ls -la $O/*.ncu-rep; tail -2 $O/ncu1.log $O/ncu2.log
fi
