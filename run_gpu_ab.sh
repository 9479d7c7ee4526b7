mkdir -p gpurun_out
R=r45
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -6) > gpurun_out/${R}_tests.log
tail -n 4 gpurun_out/${R}_tests.log
