# round 2, GPU call 1 (one GPU): tests, smoke, default bench line (with other_workloads), FP64 peak, sanitizer, launch list
set -x
mkdir -p gpurun_out
T=r2_run1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/tools/fp64_peak.cu -o /tmp/fp64_peak && timeout 120 /tmp/fp64_peak > gpurun_out/${T}_fp64_peak.jsonl 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/tools/sanitizer_smoke.py > gpurun_out/${T}_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python profiles/tools/sanitizer_smoke.py > gpurun_out/${T}_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu.log 2>&1
