set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cooperative or c4" > gpurun_out/s3_pytest_mb3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_mb3.log
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu > gpurun_out/s3_c4_mb3.log 2> gpurun_out/s3_c4_mb3.err
