mkdir -p gpurun_out
R=r79
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -k "decode" 2>&1 | tail -n 30) > gpurun_out/${R}_kernels.log
(timeout 600 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_K.json
(ICK_DECODE_TMA=0 timeout 600 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_K_notma.json
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 -k "beam or predict" 2>&1 | tail -n 5) > gpurun_out/${R}_model.log
tail -n 12 gpurun_out/${R}_kernels.log; tail -n 3 gpurun_out/${R}.err gpurun_out/${R}_model.log; cut -c1-330 gpurun_out/${R}_predict*.json
