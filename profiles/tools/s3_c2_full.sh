set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_message -s 110 -c 8 -o gpurun_out/s3_c2_final_full python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s3_ncu_c2_final_full.log 2>&1
