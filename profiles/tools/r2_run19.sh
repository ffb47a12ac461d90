# round 2, GPU call 19 (one GPU): K1 element pass with the no-evidence fast path
set -x
mkdir -p gpurun_out
T=r2_run19
timeout 900 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_generic -c 60 --csv --log-file gpurun_out/${T}_c5s_k1_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
true
