for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_n$N.err
  echo "N=$N rc=$?"; tail -c 300 gpurun_out/r2j_n$N.err
done
