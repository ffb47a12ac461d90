set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_final3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_final3_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_final3_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/s3_final3_c2.json 2> gpurun_out/s3_final3_c2.err
