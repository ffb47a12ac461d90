set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu > gpurun_out/s3_c5_default.log 2> gpurun_out/s3_c5_default.err
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --coop 8 > gpurun_out/s3_c5_coop8.log 2> gpurun_out/s3_c5_coop8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_c5_launches.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --batch 32 > gpurun_out/s3_ncu_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_tilewalk|k_iscal|k_message" -s 141 -c 47 --csv --log-file gpurun_out/s3_c3_v3_launches.csv python bench.py --workload c3 --steps 1 --no-cpu > gpurun_out/s3_ncu_c3_v3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tilewalk -s 121 -c 2 -o gpurun_out/s3_c3_tilewalk_v3_full python bench.py --workload c3 --steps 1 --no-cpu > gpurun_out/s3_ncu_c3_v3_full.log 2>&1
timeout 300 python bench.py --workload c4 --steps 3 --no-cpu > gpurun_out/s3_c4.log 2> gpurun_out/s3_c4.err
