set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_message_smem_mw -s 102 -c 1 -o gpurun_out/s3_c4_mw_full python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/s3_ncu_c4_mw_full.log 2>&1
