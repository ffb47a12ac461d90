set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_final_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/s3_final_ref_c2.json 2> gpurun_out/s3_final_ref_c2.err
timeout 600 python bench.py > gpurun_out/s3_final_c2.json 2> gpurun_out/s3_final_c2.err
timeout 600 python bench.py --e2e-batches 4 --no-cpu --steps 60 > gpurun_out/s3_final_c2_e2e4.json 2> gpurun_out/s3_final_c2_e2e4.err
timeout 600 python bench.py --workload c3 --steps 10 > gpurun_out/s3_final_c3.json 2> gpurun_out/s3_final_c3.err
timeout 600 python bench.py --workload c4 --steps 5 > gpurun_out/s3_final_c4.json 2> gpurun_out/s3_final_c4.err
timeout 900 python bench.py --workload c5 --steps 3 > gpurun_out/s3_final_c5.json 2> gpurun_out/s3_final_c5.err
timeout 600 python bench.py --workload c2s --steps 50 --no-cpu > gpurun_out/s3_final_c2s.json 2> gpurun_out/s3_final_c2s.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_message|k_generic_occ" -s 333 -c 111 --csv --log-file gpurun_out/s3_final_c4_traffic.csv python bench.py --workload c4 --steps 1 --no-cpu > gpurun_out/s3_final_ncu_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/s3_final_c4_launches.csv python bench.py --workload c4 --steps 1 --no-cpu --batch 1024 > gpurun_out/s3_final_ncu_c4b.log 2>&1
