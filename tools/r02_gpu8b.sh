#!/bin/bash
# final eight-GPU numbers (papers100M 8 and 4 GPUs, 20-step driver configuration, MAG240M)
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t0=$SECONDS
$TR --nproc-per-node 8 --master-port 29911 bench.py --gpus 8 --steps 200 --warmup 20 > $O/r02_bench_papers_g8.json 2> $O/r02_bench_papers_g8.err
echo "papers g8 rc=$? t=$((SECONDS-t0))"; head -c 260 $O/r02_bench_papers_g8.json; echo
$TR --nproc-per-node 8 --master-port 29912 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_papers_g8_s20.json 2> $O/r02_bench_papers_g8_s20.err
echo "papers g8 s20 rc=$? t=$((SECONDS-t0))"; head -c 260 $O/r02_bench_papers_g8_s20.json; echo
$TR --nproc-per-node 4 --master-port 29913 bench.py --gpus 4 --steps 200 --warmup 20 > $O/r02_bench_papers_g4.json 2> $O/r02_bench_papers_g4.err
echo "papers g4 rc=$? t=$((SECONDS-t0))"; head -c 260 $O/r02_bench_papers_g4.json; echo
$TR --nproc-per-node 8 --master-port 29914 bench.py --gpus 8 --workload mag240m --steps 200 --warmup 20 > $O/r02_bench_mag240m_g8.json 2> $O/r02_bench_mag240m_g8.err
echo "mag g8 rc=$? t=$((SECONDS-t0))"; head -c 260 $O/r02_bench_mag240m_g8.json; echo
