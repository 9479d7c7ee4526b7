mkdir -p gpurun_out
R=r90
(timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 300 -k "pickles or beam or predict" 2>&1 | tail -n 12) > gpurun_out/${R}_tests.log
tail -n 12 gpurun_out/${R}_tests.log
