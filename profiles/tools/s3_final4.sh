set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_final4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_final4_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_final4_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/s3_final4_c2.json 2> gpurun_out/s3_final4_c2.err
timeout 600 python bench.py --workload c3 --steps 10 --no-cpu > gpurun_out/s3_final4_c3.json 2> gpurun_out/s3_final4_c3.err
timeout 600 python bench.py --workload c3l --steps 10 --no-cpu > gpurun_out/s3_final4_c3l.json 2> gpurun_out/s3_final4_c3l.err
timeout 600 python bench.py --workload c4 --steps 5 --no-cpu > gpurun_out/s3_final4_c4.json 2> gpurun_out/s3_final4_c4.err
timeout 600 python bench.py --workload c2s --steps 50 --no-cpu > gpurun_out/s3_final4_c2s.json 2> gpurun_out/s3_final4_c2s.err
