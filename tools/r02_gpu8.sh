#!/bin/bash
# Round-2 eight-GPU session (gpurun --gpus 8): the north-star configuration and its reference arm.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t0=$SECONDS
$TR --nproc-per-node 8 --master-port 29901 bench.py --gpus 8 --steps 200 --warmup 20 > $O/r02_bench_papers_g8.json 2> $O/r02_bench_papers_g8.err
echo "papers g8 rc=$? t=$((SECONDS-t0))"; tail -c 1800 $O/r02_bench_papers_g8.json | cut -c1-1800; tail -2 $O/r02_bench_papers_g8.err
$TR --nproc-per-node 8 --master-port 29902 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_ref_papers_g8.json 2> $O/r02_bench_ref_papers_g8.err
echo "ref g8 rc=$? t=$((SECONDS-t0))"; tail -c 900 $O/r02_bench_ref_papers_g8.json; tail -2 $O/r02_bench_ref_papers_g8.err
$TR --nproc-per-node 4 --master-port 29903 bench.py --gpus 4 --steps 200 --warmup 20 > $O/r02_bench_papers_g4.json 2> $O/r02_bench_papers_g4.err
echo "papers g4 rc=$? t=$((SECONDS-t0))"; tail -c 700 $O/r02_bench_papers_g4.json
$TR --nproc-per-node 8 --master-port 29904 bench.py --gpus 8 --workload mag240m --steps 200 --warmup 20 > $O/r02_bench_mag240m_g8.json 2> $O/r02_bench_mag240m_g8.err
echo "mag g8 rc=$? t=$((SECONDS-t0))"; tail -c 1500 $O/r02_bench_mag240m_g8.json; tail -2 $O/r02_bench_mag240m_g8.err
$TR --nproc-per-node 8 --master-port 29905 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_papers_g8_s20.json 2> $O/r02_bench_papers_g8_s20.err
echo "papers g8 s20 rc=$? t=$((SECONDS-t0))"; head -c 300 $O/r02_bench_papers_g8_s20.json
