set -x
mkdir -p gpurun_out
for v in ch12 ch16; do
PGBP_B200_LIB=$PWD/phylogaussianbeliefprop.jl_b200/lib/libpgbp_b200_$v.so timeout 300 python bench.py --steps 100 --no-cpu > gpurun_out/s3_c2_$v.log 2> gpurun_out/s3_c2_$v.err
PGBP_B200_LIB=$PWD/phylogaussianbeliefprop.jl_b200/lib/libpgbp_b200_$v.so timeout 300 python bench.py --workload c4 --steps 4 --no-cpu > gpurun_out/s3_c4_$v.log 2> gpurun_out/s3_c4_$v.err
done
