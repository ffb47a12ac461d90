# round 2, GPU call 2 (two GPUs): the fused NVLink peer-window gather against the asynchronous NCCL all-gather
set -x
mkdir -p gpurun_out
T=r2_run2
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
for G in peer nccl; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-others --gather $G > gpurun_out/${T}_c2_${G}.json 2> gpurun_out/${T}_c2_${G}.err; echo "rc=$?" >> gpurun_out/${T}_c2_${G}.err
done
# the same on one GPU right before / after (same box): the N = 1 denominator
timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu > gpurun_out/${T}_c2_1gpu.json 2> gpurun_out/${T}_c2_1gpu.err
# full default line on two GPUs (other_workloads under torchrun)
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_default_2gpu.json 2> gpurun_out/${T}_default_2gpu.err; echo "rc=$?" >> gpurun_out/${T}_default_2gpu.err
