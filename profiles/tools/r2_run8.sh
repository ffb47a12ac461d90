# round 2, GPU call 8 (eight GPUs of one box): the default bench line under torchrun, as the driver launches it
set -x
mkdir -p gpurun_out
T=r2_run8
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
free -g > gpurun_out/${T}_hostmem.txt 2>&1; nproc >> gpurun_out/${T}_hostmem.txt
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${T}_default_8gpu.json 2> gpurun_out/${T}_default_8gpu.err; echo "rc=$?" >> gpurun_out/${T}_default_8gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu > gpurun_out/${T}_c2_1gpu.json 2> gpurun_out/${T}_c2_1gpu.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
