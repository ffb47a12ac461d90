set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lazaridis or c2 or canonicalform or autostop or failed" > gpurun_out/s3_pytest_pin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_pin.log
timeout 600 python bench.py --steps 100 --no-cpu > gpurun_out/s3_c2_pin.log 2> gpurun_out/s3_c2_pin.err
timeout 600 python bench.py --steps 100 --no-cpu --e2e-batches 4 > gpurun_out/s3_c2_pin4.log 2> gpurun_out/s3_c2_pin4.err
