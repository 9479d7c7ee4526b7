mkdir -p gpurun_out
R=r51
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -5) > gpurun_out/${R}_tests.log
tail -n 5 gpurun_out/${R}_tests.log
