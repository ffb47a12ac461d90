set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c5 or multivariate or assignfactors" > gpurun_out/s3_pytest_k1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_k1.log
timeout 900 python bench.py --workload c5 --steps 3 > gpurun_out/s3_bench_c5.json 2> gpurun_out/s3_bench_c5.err
