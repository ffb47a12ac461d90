# round 2, GPU call 18 (two GPUs of one box): per-device kernel attributes -- the whole GPU suite including the
# two-devices-in-one-process test, then a short C2 + C4 bench for regressions
set -x
mkdir -p gpurun_out
T=r2_run18
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-cpu --no-c5 > gpurun_out/${T}_default_noc5.json 2> gpurun_out/${T}_default_noc5.err; echo "rc=$?" >> gpurun_out/${T}_default_noc5.err
