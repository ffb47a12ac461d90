# round 2, GPU call 3 (one GPU): factored shared-precision path (k_jmsg + k_hmsg): parity on the GPU, c2s and c5s bench lines
set -x
mkdir -p gpurun_out
T=r2_run3
timeout 900 python -m pytest tests -m gpu -x -q -k "shared or multirank or abi" > gpurun_out/${T}_pytest_shared.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_shared.log
timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others > gpurun_out/${T}_c2s.json 2> gpurun_out/${T}_c2s.err; echo "rc=$?" >> gpurun_out/${T}_c2s.err
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 4 > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --batch 256 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_c2s_launches.csv python bench.py --workload c2s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c2s.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_all.log
