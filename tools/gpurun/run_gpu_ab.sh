mkdir -p gpurun_out
R=r53
(timeout 600 python -m pytest tests -m gpu -q --tb=short -x --timeout 300 -k "pool_rows or encoder_head" 2>&1 | tail -4) > gpurun_out/${R}_kernels.log
(timeout 300 python tools/microbench.py 2>&1 | tail -3) > gpurun_out/${R}_micro.log
cat gpurun_out/${R}_kernels.log gpurun_out/${R}_micro.log
