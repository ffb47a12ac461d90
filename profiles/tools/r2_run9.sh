# round 2, GPU call 9 (one GPU): element pass of the shared path -- bulk-copy staged kernel (k_hmsg_bulk) against the
# thread-per-element kernel at 3 / 4 / 5 resident blocks per SM
set -x
mkdir -p gpurun_out
T=r2_run9
L=phylogaussianbeliefprop.jl_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_bulk.json 2> gpurun_out/${T}_c5s_bulk.err; echo "rc=$?" >> gpurun_out/${T}_c5s_bulk.err
PGBP_HMSG_BULK=0 timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_mb4.json 2> gpurun_out/${T}_c5s_mb4.err; echo "rc=$?" >> gpurun_out/${T}_c5s_mb4.err
for mb in 3 5; do
PGBP_B200_LIB=$PWD/$L/libpgbp_b200_mb$mb.so timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_mb$mb.json 2> gpurun_out/${T}_c5s_mb$mb.err; echo "rc=$?" >> gpurun_out/${T}_c5s_mb$mb.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
bash profiles/tools/ncu_full.sh ${T}_hmsg_bulk 'k_hmsg_bulk' 40 2 --workload c5s --batch 512 --steps 1 --warmup 1 --no-cpu --no-others
