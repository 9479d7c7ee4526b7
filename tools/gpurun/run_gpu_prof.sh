mkdir -p gpurun_out
R=${R:-r03j}
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fwd_pkernel|wgrad_group_tc_kernel" -s 0 -c 8 -o gpurun_out/${R}_fwd_wgrad python tools/step_prof.py 2 > gpurun_out/${R}_ncu.log 2>&1
tail -n 2 gpurun_out/${R}_ncu.log; ls -la gpurun_out/${R}_fwd_wgrad.ncu-rep
