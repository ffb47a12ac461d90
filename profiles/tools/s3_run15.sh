set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_message_smem_mwp -s 100 -c 2 -o gpurun_out/s3_c5_mwp_full python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --batch 32 > gpurun_out/s3_ncu_c5_mwp_full.log 2>&1
