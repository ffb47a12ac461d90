# round 2, GPU call 22 (one GPU): DRAM bytes of the message kernels of C5S, per launch (roofline.traffic of that line)
set -x
mkdir -p gpurun_out
T=r2_run22
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_hmsg|k_jmsg|k_hwalk|k_jwalk' -c 1200 --csv --log-file gpurun_out/${T}_c5s_traffic_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu.log 2>&1
true
