set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "assignfactors or synth or c4 or lazaridis or lazy" > gpurun_out/s3_pytest_k1b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_k1b.log
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu > gpurun_out/s3_c4_k1b.log 2> gpurun_out/s3_c4_k1b.err
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu > gpurun_out/s3_c5_k1b.log 2> gpurun_out/s3_c5_k1b.err
