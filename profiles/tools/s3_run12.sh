set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cooperative or c5 or multivariate" > gpurun_out/s3_pytest_tile.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_tile.log
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu > gpurun_out/s3_c5_tile.log 2> gpurun_out/s3_c5_tile.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_c5_launches_tile.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --batch 32 > gpurun_out/s3_ncu_c5_tile.log 2>&1
