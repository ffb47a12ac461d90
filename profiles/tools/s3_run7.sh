set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cooperative or c5 or multivariate" > gpurun_out/s3_pytest_mw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_mw.log
for md in -1 2 1; do
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu --coop $md > gpurun_out/s3_c5_mw_$md.log 2> gpurun_out/s3_c5_mw_$md.err
done
