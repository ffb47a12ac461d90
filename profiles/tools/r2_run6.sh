# round 2, GPU call 6 (one GPU): row-parallel launcher for the group batch's K1, --set full captures of the shared-path kernels
set -x
mkdir -p gpurun_out
T=r2_run6
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_generic|k_aos" -c 60 --csv --log-file gpurun_out/${T}_c5s_k1_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_generic_occ" -s 9 -c 2 -o gpurun_out/${T}_k1 python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hmsg" -s 4 -c 3 -o gpurun_out/${T}_hmsg python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_jmsg" -s 1 -c 3 -o gpurun_out/${T}_jmsg python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full2.log 2>&1
