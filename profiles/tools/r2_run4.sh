# round 2, GPU call 4 (one GPU): shared-precision path v2 (group pass on its own stream, multi-warp k_jmsg, K1 block cache,
# element-pass chunk 4 vs 8), C2 pipeline-chunk sweep
set -x
mkdir -p gpurun_out
T=r2_run4
timeout 600 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest_shared.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_shared.log
for CH in 8 4; do
  PGBP_HMSG_CHUNK=$CH timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5s_ch${CH}.json 2> gpurun_out/${T}_c5s_ch${CH}.err; echo "rc=$?" >> gpurun_out/${T}_c5s_ch${CH}.err
  PGBP_HMSG_CHUNK=$CH timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others --cpu-seconds 2 > gpurun_out/${T}_c2s_ch${CH}.json 2> gpurun_out/${T}_c2s_ch${CH}.err; echo "rc=$?" >> gpurun_out/${T}_c2s_ch${CH}.err
done
for PL in 2 6 8 16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu --pipeline $PL > gpurun_out/${T}_c2_pipe${PL}.json 2> gpurun_out/${T}_c2_pipe${PL}.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
