# round 2, GPU call 10 (one GPU): group pass of the shared path on wide levels (one warp per message instead of a
# 128-thread block), K1 element pass looked at line by line
set -x
mkdir -p gpurun_out
T=r2_run10
timeout 900 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
PGBP_JMSG_WIDE=1000000000 timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_jblock.json 2> gpurun_out/${T}_c5s_jblock.err; echo "rc=$?" >> gpurun_out/${T}_c5s_jblock.err
timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others --no-cpu > gpurun_out/${T}_c2s.json 2> gpurun_out/${T}_c2s.err; echo "rc=$?" >> gpurun_out/${T}_c2s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
NCU_TOP=24 bash profiles/tools/ncu_full.sh ${T}_k1h 'k_generic_occ' 1 1 --workload c5s --batch 512 --steps 1 --warmup 1 --no-cpu --no-others
