TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2 1; do
  $TR --nproc-per-node $n --master-port $((29710 + n)) bench.py --gpus $n --workload corpus --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_corpus20_n$n.json 2> gpurun_out/r02_corpus20_n$n.err
done
