mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1_gpu.txt 2>&1
(timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "not True" --timeout 600 2>&1 | tail -80) > gpurun_out/r1_kernels_simt.log
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "True" --timeout 300 2>&1 | tail -80) > gpurun_out/r1_kernels_tc.log
(ICKB200_NO_TC=1 timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -80) > gpurun_out/r1_model_notc.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -80) > gpurun_out/r1_model_tc.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -20) > gpurun_out/r1_smoke.log
tail -5 gpurun_out/r1_*.log
