mkdir -p gpurun_out
R=r29
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -4) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 220 -c 240 --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 2 > gpurun_out/${R}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wgrad_group_tc_kernel|gemm_tn_tc_kernel|bwd_dkv_pkernel|bwd_dq_pkernel" -s 100 -c 14 -o gpurun_out/${R}_top -f python tools/step_prof.py 2 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 3 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_bench.json; tail -n 2 gpurun_out/${R}_ncu2.log
