set -x
mkdir -p gpurun_out
for wd in 128 256 512; do
  timeout 300 python bench.py --workload c3 --steps 5 --no-cpu --tw-wide $wd > gpurun_out/s3_c3_wide_$wd.log 2> gpurun_out/s3_c3_wide_$wd.err
done
timeout 300 python profiles/tools/e2e_probe.py > gpurun_out/s3_e2e_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_tilewalk|k_iscal|k_message" -s 141 -c 47 --csv --log-file gpurun_out/s3_c3_launches.csv python bench.py --workload c3 --steps 1 --no-cpu > gpurun_out/s3_ncu_c3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tilewalk -s 121 -c 2 -o gpurun_out/s3_c3_tilewalk_full python bench.py --workload c3 --steps 1 --no-cpu > gpurun_out/s3_ncu_c3_full.log 2>&1
ls -la gpurun_out
