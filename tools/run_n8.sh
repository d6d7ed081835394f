#!/bin/bash
# 8-GPU measurements of round 2 (one box): contract line at N=8 and N=2, corpus strong scaling at 1/2/4/8, train_step at 8.
set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
$TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29610 + n)) bench.py --gpus $n --workload corpus --steps 3 --no-cpu > gpurun_out/r02_corpus_n$n.json 2> gpurun_out/r02_corpus_n$n.err
done
$TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --workload train_step --steps 30 > gpurun_out/r02_train_n8.json 2> gpurun_out/r02_train_n8.err
python bench.py --workload train_step --steps 30 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
lscpu | head -25 > gpurun_out/r02_lscpu.txt 2>&1
echo done
