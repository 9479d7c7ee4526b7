mkdir -p gpurun_out
R=r46
(timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=line -x --timeout 1100 -k "wgrad_group or gemm_dual or add_ln_dual or (mha and bfloat16 and 301)" 2>&1 | tail -30) > gpurun_out/${R}_memcheck.log; echo "rc=$?"
tail -n 30 gpurun_out/${R}_memcheck.log
