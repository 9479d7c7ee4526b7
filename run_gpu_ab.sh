mkdir -p gpurun_out
R=r28
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "wgrad" 2>&1 | tail -15) > gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -8) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
cat gpurun_out/${R}_kernels.log; tail -3 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_bench.json; tail -5 gpurun_out/${R}_bench.err
