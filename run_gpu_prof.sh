mkdir -p gpurun_out
R=r47
(timeout 600 python bench.py --steps 50 --warmup 3 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/${R}_ref.err | tail -1) > gpurun_out/${R}_ref.json
(timeout 300 python tools/microbench.py 2>&1) > gpurun_out/${R}_micro.log
python tools/step_prof.py 3 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "profile_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 3 > gpurun_out/${R}_ncu1.log 2>&1
ncu --nvtx --nvtx-include "profile_step/" --set full --clock-control none --import-source on -k regex:"wgrad_group_tc_kernel|gemm_tn_tc_kernel|bwd_dkv_pkernel|bwd_dq_pkernel|bwd_dq_ds_kernel|fwd_pkernel|add_ln_bwd_fast" -s 60 -c 16 -o gpurun_out/${R}_top -f python tools/step_prof.py 3 > gpurun_out/${R}_ncu2.log 2>&1
cut -c1-200 gpurun_out/${R}_bench.json; cut -c1-200 gpurun_out/${R}_ref.json; tail -n 2 gpurun_out/${R}_ncu1.log gpurun_out/${R}_ncu2.log; grep -c gpu__time_duration gpurun_out/${R}_launches.csv
