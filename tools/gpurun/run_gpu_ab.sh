mkdir -p gpurun_out
R=${R:-r02x}
(timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "gemm or wgrad or pool_rows" 2>&1 | tail -n 15) > gpurun_out/${R}_kernels.log
tail -n 6 gpurun_out/${R}_kernels.log
(timeout 600 python bench.py --steps 30 --warmup 3 --no-decode 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
tail -n 3 gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench.json
