set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_gpu3.log
timeout 900 python bench.py --workload c5 --steps 3 > gpurun_out/s3_bench_c5.json 2> gpurun_out/s3_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_c5_launches_mw.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu --batch 32 > gpurun_out/s3_ncu_c5_mw.log 2>&1
timeout 600 python bench.py --workload c2s --steps 50 --no-cpu > gpurun_out/s3_bench_c2s.json 2> gpurun_out/s3_bench_c2s.err
