set -x
mkdir -p gpurun_out
for md in -1 3 5; do
timeout 600 python bench.py --workload c4 --steps 4 --no-cpu --coop $md > gpurun_out/s3_c4_coop_$md.log 2> gpurun_out/s3_c4_coop_$md.err
done
