mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r3_launches.csv $CMD > gpurun_out/r3_ncu1.log 2>&1
$CMD > gpurun_out/r3_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mha_bwd -s 24 -c 4 -o gpurun_out/r3_mha_bwd $CMD > gpurun_out/r3_ncu2.log 2>&1
$CMD > gpurun_out/r3_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 100 -c 3 -o gpurun_out/r3_gemm_tc $CMD > gpurun_out/r3_ncu3.log 2>&1
ls -la gpurun_out/ | tail -20; tail -3 gpurun_out/r3_ncu*.log
