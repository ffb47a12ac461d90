set -x
mkdir -p gpurun_out
N=${1:-8}
for wl in c2 c3 c4; do
  st=20; [ $wl = c2 ] && st=100; [ $wl = c4 ] && st=5
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps $st --warmup 3 > gpurun_out/s3_bench_${wl}_n$N.json 2> gpurun_out/s3_bench_${wl}_n$N.err
done
