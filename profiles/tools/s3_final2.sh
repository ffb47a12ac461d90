set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_final2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_final2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_final2_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/s3_final2_c2.json 2> gpurun_out/s3_final2_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3_final2_c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s3_final2_ncu_c2.log 2>&1
