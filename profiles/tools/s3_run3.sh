set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_gpu.log
for ln in 8 16 4; do
timeout 300 python bench.py --workload c3 --steps 5 --no-cpu --tw-lanes $ln > gpurun_out/s3_c3_v3_l$ln.log 2> gpurun_out/s3_c3_v3_l$ln.err
done
for nb in 2 3; do
  timeout 300 python bench.py --steps 100 --no-cpu --e2e-batches $nb > gpurun_out/s3_c2_lazy_e2e_$nb.log 2> gpurun_out/s3_c2_lazy_e2e_$nb.err
done
timeout 300 python profiles/tools/e2e_probe.py > gpurun_out/s3_e2e_probe2.log 2>&1
