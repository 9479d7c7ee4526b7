mkdir -p gpurun_out
R=r38
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "add_ln" 2>&1 | tail -12) > gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -6) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
cat gpurun_out/${R}_kernels.log; tail -n 4 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_bench.json; tail -n 3 gpurun_out/${R}_bench.err
