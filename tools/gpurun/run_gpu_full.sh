mkdir -p gpurun_out
R=r92
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 280 -k "decode" 2>&1 | tail -n 8) > gpurun_out/${R}_kernels.log
tail -n 8 gpurun_out/${R}_kernels.log
