mkdir -p gpurun_out
R=${R:-r03f}
(timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "pointer or indicators or grouped" 2>&1 | tail -n 15) > gpurun_out/${R}_kernels.log
tail -n 6 gpurun_out/${R}_kernels.log
(timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_baseline_parity.py -m gpu -q --tb=short -x --timeout 500 2>&1 | tail -n 8) > gpurun_out/${R}_model.log
tail -n 4 gpurun_out/${R}_model.log
(timeout 600 python bench.py --steps 30 --warmup 3 --no-decode 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
tail -n 3 gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench.json
