set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lazy or c2 or c3 or pipelined or loopy" > gpurun_out/s3_pytest_reset.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_reset.log
timeout 600 python bench.py > gpurun_out/s3_bench_c2_b.json 2> gpurun_out/s3_bench_c2_b.err
timeout 600 python bench.py --workload c3 --steps 10 --no-cpu > gpurun_out/s3_bench_c3_b.json 2> gpurun_out/s3_bench_c3_b.err
