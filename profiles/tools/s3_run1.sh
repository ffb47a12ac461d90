set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tilewalk or c3" > gpurun_out/s3_pytest_tw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_tw.log
for cfg in "0 0 0" "-1 0 0" "-1 16 0" "-1 4 0" "-1 8 1000" "-1 8 24" "-1 16 128"; do
  set -- $cfg
  timeout 300 python bench.py --workload c3 --steps 5 --no-cpu --tilewalk $1 --tw-lanes $2 --tw-wide $3 > gpurun_out/s3_c3_tw_$1_$2_$3.log 2> gpurun_out/s3_c3_tw_$1_$2_$3.err
done
for nb in 2 3 4; do
  timeout 300 python bench.py --steps 100 --no-cpu --e2e-batches $nb > gpurun_out/s3_c2_e2e_$nb.log 2> gpurun_out/s3_c2_e2e_$nb.err
done
