mkdir -p gpurun_out
python tools/step_prof.py 2 > gpurun_out/r15_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 420 --csv --log-file gpurun_out/r15_launches.csv python tools/step_prof.py 2 > gpurun_out/r15_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_tc_kernel|wgrad_tc_kernel|add_ln" -s 150 -c 24 -o gpurun_out/r15_gemm -f python tools/step_prof.py 2 > gpurun_out/r15_ncu2.log 2>&1
tail -3 gpurun_out/r15_plain.log gpurun_out/r15_ncu1.log gpurun_out/r15_ncu2.log
