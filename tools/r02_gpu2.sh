#!/bin/bash
# Round-2 two-GPU measurement session (gpurun --gpus 2); everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
what="${*:-tests bench ab nccl ncu}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
if [[ $what == *tests* ]]; then
  (time python -m pytest tests/test_gpu_multi.py -x -q) > $O/r02_tests_g2.log 2>&1
  echo "pytest rc=$?" >> $O/r02_tests_g2.log
  tail -5 $O/r02_tests_g2.log
fi
if [[ $what == *bench* ]]; then
  $TR --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 > $O/r02_bench_papers_g2.json 2> $O/r02_bench_papers_g2.err
  echo "bench rc=$?"; tail -c 2500 $O/r02_bench_papers_g2.json; tail -3 $O/r02_bench_papers_g2.err
  $TR --master-port 29612 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_ref_papers_g2.json 2> $O/r02_bench_ref_papers_g2.err
  echo "ref rc=$?"; tail -c 1200 $O/r02_bench_ref_papers_g2.json; tail -3 $O/r02_bench_ref_papers_g2.err
  $TR --master-port 29613 bench.py --gpus 2 --workload products --steps 200 --warmup 20 > $O/r02_bench_products_g2.json 2> $O/r02_bench_products_g2.err
  echo "products rc=$?"; tail -c 1500 $O/r02_bench_products_g2.json
  $TR --master-port 29614 bench.py --gpus 2 --workload products-layerwise --steps 200 --warmup 20 > $O/r02_bench_products_layerwise_g2.json 2> $O/r02_bench_products_layerwise_g2.err
  echo "layerwise rc=$?"; tail -c 800 $O/r02_bench_products_layerwise_g2.json; tail -3 $O/r02_bench_products_layerwise_g2.err
fi
if [[ $what == *ab* ]]; then
  $TR --master-port 29615 tools/gather_ab.py > $O/r02_ab_gather_bulk_g2.txt 2> $O/r02_ab_gather_bulk_g2.err
  cat $O/r02_ab_gather_bulk_g2.txt; tail -3 $O/r02_ab_gather_bulk_g2.err
fi
if [[ $what == *nccl* ]]; then
  $TR --master-port 29616 tools/p2p_check.py > $O/r02_p2p_vs_nccl_g2.txt 2> $O/r02_p2p_vs_nccl_g2.err
  cat $O/r02_p2p_vs_nccl_g2.txt; tail -3 $O/r02_p2p_vs_nccl_g2.err
fi
if [[ $what == *ncu* ]]; then
  M=nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes.sum.per_second,nvltx__bytes.sum.per_second,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
  for rb in 256 200 1536; do
    python tools/ncu_peer_gather.py --row-bytes $rb > $O/r02_peer_gather_plain_$rb.json 2>&1 &&
    ncu --metrics $M --clock-control none -k regex:k_gather -s 3 -c 2 --csv --log-file $O/r02_ncu_nvlink_peer_gather_$rb.csv \
        python tools/ncu_peer_gather.py --row-bytes $rb > $O/ncu_peer_$rb.log 2>&1
    cat $O/r02_peer_gather_plain_$rb.json; grep -v "^==" $O/r02_ncu_nvlink_peer_gather_$rb.csv | cut -d, -f5,13,14,15 | head -16
  done
  python tools/ncu_peer_gather.py --row-bytes 256 --bulk 1 > $O/r02_peer_gather_plain_256_bulk.json 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:k_gather -s 3 -c 2 --csv --log-file $O/r02_ncu_nvlink_peer_gather_256_bulk.csv \
      python tools/ncu_peer_gather.py --row-bytes 256 --bulk 1 > $O/ncu_peer_256_bulk.log 2>&1
  cat $O/r02_peer_gather_plain_256_bulk.json; grep -v "^==" $O/r02_ncu_nvlink_peer_gather_256_bulk.csv | cut -d, -f5,13,14,15 | head -16
fi
