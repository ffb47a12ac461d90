# round 2, GPU call 5 (one GPU): bulk-copy (TMA 1-D) staging in the multi-warp kernels, shared path v3 (lazy sepset zero,
# records staged in shared memory), full GPU test suite, --set full captures of the new kernels
set -x
mkdir -p gpurun_out
T=r2_run5
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; echo "rc=$?" >> gpurun_out/${T}_c5.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err; echo "rc=$?" >> gpurun_out/${T}_c4.err
timeout 900 python bench.py --workload c5s --steps 3 --warmup 3 --no-others --cpu-seconds 2 > gpurun_out/${T}_c5s.json 2> gpurun_out/${T}_c5s.err; echo "rc=$?" >> gpurun_out/${T}_c5s.err
timeout 600 python bench.py --workload c2s --steps 20 --warmup 5 --no-others --cpu-seconds 2 > gpurun_out/${T}_c2s.json 2> gpurun_out/${T}_c2s.err; echo "rc=$?" >> gpurun_out/${T}_c2s.err
# launch list of one full-size c5s step without the message kernels (what is the rest of the step made of?)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"Assign|aos|soa|fill|Theta|integrate|iscal" -c 200 --csv --log-file gpurun_out/${T}_c5s_k1_launches.csv python bench.py --workload c5s --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_c5s_k1.log 2>&1
# --set full: the element-pass and group-pass kernels on wide launches (c5s, batch 512), the bulk-staged multi-warp kernels (c5, batch 32)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hmsg<16" -s 40 -c 2 -o gpurun_out/${T}_hmsg16 python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_jmsg<128" -s 2 -c 2 -o gpurun_out/${T}_jmsg128 python bench.py --workload c5s --batch 512 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_message_smem_mw" -s 6 -c 3 -o gpurun_out/${T}_mw python bench.py --workload c5 --batch 32 --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/${T}_ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/${T}_ncu_full3.log 2>&1
