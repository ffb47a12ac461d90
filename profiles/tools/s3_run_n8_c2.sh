set -x
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c2 --steps 100 --warmup 3 > gpurun_out/s3_final_c2_n8.json 2> gpurun_out/s3_final_c2_n8.err
