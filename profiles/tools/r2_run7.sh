# round 2, GPU call 7 (one GPU): walk kernels for narrow runs of the shared path, lean K1 element pass, staged records via LDS
set -x
mkdir -p gpurun_out
T=r2_run7
timeout 900 python -m pytest tests -m gpu -x -q -k "shared or adjud or fullsize" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5s_walk.json 2> gpurun_out/${T}_c5s_walk.err; echo "rc=$?" >> gpurun_out/${T}_c5s_walk.err
PGBP_SHARED_WALK=0 timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s_nowalk.json 2> gpurun_out/${T}_c5s_nowalk.err; echo "rc=$?" >> gpurun_out/${T}_c5s_nowalk.err
timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others --cpu-seconds 2 > gpurun_out/${T}_c2s.json 2> gpurun_out/${T}_c2s.err; echo "rc=$?" >> gpurun_out/${T}_c2s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
