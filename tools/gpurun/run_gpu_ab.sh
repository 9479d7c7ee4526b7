mkdir -p gpurun_out
R=r94
(timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 150 -k "beam_select" 2>&1 | tail -n 6) > gpurun_out/${R}_kernels.log
(timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 250 -k "beam" 2>&1 | tail -n 6) > gpurun_out/${R}_model.log
(timeout 200 python tools/bench_predict.py --variant K --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_beam_K.json
(ICK_BEAM_TOPK=scan timeout 200 python tools/bench_predict.py --variant K --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_beam_K_scan.json
tail -n 4 gpurun_out/${R}_kernels.log gpurun_out/${R}_model.log gpurun_out/${R}.err; cut -c1-200 gpurun_out/${R}_beam_K.json gpurun_out/${R}_beam_K_scan.json
