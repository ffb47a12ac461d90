# round 2, GPU call 15 (N GPUs of one box, N = $1): the default bench line of the final build under torchrun, as the
# driver launches it, plus the one-GPU line on the same box for the ratio
set -x
N=$1
mkdir -p gpurun_out
T=r2_run15_n$N
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err; echo "rc=$?" >> gpurun_out/${T}_default.err
