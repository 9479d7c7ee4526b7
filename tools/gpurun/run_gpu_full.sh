mkdir -p gpurun_out
R=r96
(timeout 120 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 100 -k "validate_step" 2>&1 | tail -n 8) > gpurun_out/${R}_tests.log
tail -n 8 gpurun_out/${R}_tests.log
