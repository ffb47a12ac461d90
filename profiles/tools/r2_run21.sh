# round 2, GPU call 21 (one GPU): group pass with the divide / multiply operands fetched up front
set -x
mkdir -p gpurun_out
T=r2_run21
timeout 900 python -m pytest tests -m gpu -x -q -k "shared" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others --no-cpu > gpurun_out/${T}_c2s.json 2> gpurun_out/${T}_c2s.err; echo "rc=$?" >> gpurun_out/${T}_c2s.err
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --no-cpu > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
