set -x
mkdir -p gpurun_out
for pl in 2 4; do
timeout 600 python bench.py --workload c4 --steps 4 --no-cpu --pipeline $pl > gpurun_out/s3_c4_pl_$pl.log 2> gpurun_out/s3_c4_pl_$pl.err
done
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu --pipeline 2 > gpurun_out/s3_c5_pl_2.log 2> gpurun_out/s3_c5_pl_2.err
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu --pipeline 4 > gpurun_out/s3_c5_pl_4.log 2> gpurun_out/s3_c5_pl_4.err
