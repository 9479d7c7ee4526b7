mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "gemm or wgrad" 2>&1 | tail -5) > gpurun_out/r22_kernels.log
(timeout 600 python tools/microbench.py 2>&1 | head -25) > gpurun_out/r22_micro.log
(timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r22_bench.err | tail -1) > gpurun_out/r22_bench.json
cat gpurun_out/r22_kernels.log gpurun_out/r22_micro.log; cut -c1-200 gpurun_out/r22_bench.json
