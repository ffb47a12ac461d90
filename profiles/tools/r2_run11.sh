# round 2, GPU call 11 (one GPU): dense pitch of the group batch (one group: 8 instead of 32 bytes per J entry),
# tip fast path of the K1 element pass (no spills)
set -x
mkdir -p gpurun_out
T=r2_run11
timeout 900 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
PGBP_JMSG_WIDE=1000000000 timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_jblock.json 2> gpurun_out/${T}_c5s_jblock.err; echo "rc=$?" >> gpurun_out/${T}_c5s_jblock.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
