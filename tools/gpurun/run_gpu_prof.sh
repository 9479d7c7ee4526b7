mkdir -p gpurun_out
R=${R:-r03b}
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"pointer_bwd_mma|ce_kernel|adam_kernel_v4|add_ln_bwd_fast" -s 0 -c 40 -o gpurun_out/${R}_misc python tools/step_prof.py 2 > gpurun_out/${R}_ncu.log 2>&1
tail -n 2 gpurun_out/${R}_ncu.log; ls -la gpurun_out/${R}_misc.ncu-rep
