# round 2, GPU call 13 (one GPU): verification of the final build -- the whole GPU test suite, smoke(), the default
# bench line with its other_workloads block, the reference arm, launch lists and one --set full capture
set -x
mkdir -p gpurun_out
T=r2_run13
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
( time timeout 1200 python bench.py > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err ) 2> gpurun_out/${T}_default.time; echo "rc=$?" >> gpurun_out/${T}_default.err
timeout 300 python bench.py --impl reference > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_c5s_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s.log 2>&1
NCU_TOP=12 bash profiles/tools/ncu_full.sh ${T}_hmsg_bulk_wide 'k_hmsg_bulk' 0 2 --workload c5s --batch 512 --steps 1 --warmup 1 --no-cpu --no-others
