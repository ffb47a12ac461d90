# round 2, GPU call 6b (one GPU): row-parallel launcher for the group batch's K1, --set full captures of the shared-path
# kernels summarised on the box
set -x
mkdir -p gpurun_out
T=r2_run6
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_generic|k_aos" -c 60 --csv --log-file gpurun_out/${T}_c5s_k1_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s_k1.log 2>&1
bash profiles/tools/ncu_full.sh ${T}_k1h "k_generic_occ" 8 1 --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others
bash profiles/tools/ncu_full.sh ${T}_hmsg "k_hmsg" 4 2 --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others
bash profiles/tools/ncu_full.sh ${T}_jmsg "k_jmsg" 1 2 --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others
du -sh gpurun_out
