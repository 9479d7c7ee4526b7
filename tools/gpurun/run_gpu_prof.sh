mkdir -p gpurun_out
R=r49
ncu --nvtx --nvtx-include "profile_step/" --set full --clock-control none --import-source on -k regex:"pointer_bwd_mma|pointer_fwd_mma|adam_kernel|ce_kernel" -c 6 -o gpurun_out/${R}_misc -f python tools/step_prof.py 3 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 2 gpurun_out/${R}_ncu2.log
