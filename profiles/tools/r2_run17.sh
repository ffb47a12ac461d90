# round 2, GPU call 17 (one GPU): the final build once more -- whole GPU test suite, smoke(), default bench line, reference arm
set -x
mkdir -p gpurun_out
T=r2_run17
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
( time timeout 1200 python bench.py > gpurun_out/${T}_default.json 2> gpurun_out/${T}_default.err ) 2> gpurun_out/${T}_default.time; echo "rc=$?" >> gpurun_out/${T}_default.err
timeout 300 python bench.py --impl reference > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
