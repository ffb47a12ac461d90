set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ltrip or c3" > gpurun_out/s3_pytest_ltrip.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_ltrip.log
timeout 600 python bench.py --workload c3l --steps 10 > gpurun_out/s3_bench_c3l.json 2> gpurun_out/s3_bench_c3l.err
timeout 600 python bench.py --workload c3l --steps 5 --no-cpu --tilewalk 0 > gpurun_out/s3_bench_c3l_tw0.json 2> gpurun_out/s3_bench_c3l_tw0.err
