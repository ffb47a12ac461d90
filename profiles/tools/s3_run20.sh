set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cooperative or c5 or c4 or multivariate or synth" > gpurun_out/s3_pytest_pb1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_pb1.log
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu > gpurun_out/s3_c4_pb1.log 2> gpurun_out/s3_c4_pb1.err
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu > gpurun_out/s3_c5_pb1.log 2> gpurun_out/s3_c5_pb1.err
