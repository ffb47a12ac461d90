# round 2, GPU call 14 (one GPU): launch list of C2S (where do its 0.42 ms go), --set full of the wide bulk element
# pass and of the K1 element pass after the tip fast path
set -x
mkdir -p gpurun_out
T=r2_run14
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_c2s_launches.csv python bench.py --workload c2s --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c2s.log 2>&1
NCU_TOP=12 bash profiles/tools/ncu_full.sh ${T}_hmsg_bulk_wide 'k_hmsg_bulk' 0 2 --workload c5s --batch 512 --steps 1 --warmup 1 --no-cpu --no-others
NCU_TOP=12 bash profiles/tools/ncu_full.sh ${T}_k1h 'k_generic_occ' 1 1 --workload c5s --batch 512 --steps 1 --warmup 1 --no-cpu --no-others
