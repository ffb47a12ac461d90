set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_gpu2.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/s3_bench_c2.json 2> gpurun_out/s3_bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/s3_bench_ref_c2.json 2> gpurun_out/s3_bench_ref_c2.err
timeout 600 python bench.py --workload c3 --steps 10 > gpurun_out/s3_bench_c3.json 2> gpurun_out/s3_bench_c3.err
timeout 600 python bench.py --workload c4 --steps 5 > gpurun_out/s3_bench_c4.json 2> gpurun_out/s3_bench_c4.err
