set -x
mkdir -p gpurun_out
PGBP_E2E_LOCK=1 timeout 300 python bench.py --steps 60 --no-cpu --e2e-batches 3 > gpurun_out/s3_c2_lock3.log 2> gpurun_out/s3_c2_lock3.err
PGBP_E2E_LOCK=1 timeout 300 python bench.py --steps 60 --no-cpu --e2e-batches 2 > gpurun_out/s3_c2_lock2.log 2> gpurun_out/s3_c2_lock2.err
timeout 300 python bench.py --steps 60 --no-cpu --e2e-batches 3 --pipeline 1 > gpurun_out/s3_c2_pl1.log 2> gpurun_out/s3_c2_pl1.err
