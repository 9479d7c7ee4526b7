mkdir -p gpurun_out
R=r89
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --timeout 900 2>&1 | tail -n 8) > gpurun_out/${R}_tests.log
tail -n 4 gpurun_out/${R}_tests.log
