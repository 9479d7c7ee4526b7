mkdir -p gpurun_out
R=r26
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"bwd_dkv_pkernel|bwd_dq_pkernel|fwd_pkernel" -s 36 -c 3 -o gpurun_out/${R}_attn -f python tools/step_prof.py 2 > gpurun_out/${R}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|gemm_tn_tc_kernel" -s 150 -c 12 -o gpurun_out/${R}_gemm -f python tools/step_prof.py 2 > gpurun_out/${R}_ncu3.log 2>&1
tail -2 gpurun_out/${R}_ncu2.log gpurun_out/${R}_ncu3.log
