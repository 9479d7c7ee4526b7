mkdir -p gpurun_out
R=r74
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -k "beam or grouped or indicators or greedy or pointer" 2>&1 | tail -40) > gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 -k "beam or predict" 2>&1 | tail -60) > gpurun_out/${R}_model.log
(timeout 600 python tools/bench_predict.py --variant K --beam 5 2> gpurun_out/${R}_beam.err | tail -1) > gpurun_out/${R}_beam_K.json
(timeout 600 python tools/bench_predict.py --variant K 2>> gpurun_out/${R}_beam.err | tail -1) > gpurun_out/${R}_predict_K.json
(ICK_PTR_DECODE_MMA=0 timeout 600 python tools/bench_predict.py --variant K 2>> gpurun_out/${R}_beam.err | tail -1) > gpurun_out/${R}_predict_K_simtptr.json
tail -15 gpurun_out/${R}_kernels.log gpurun_out/${R}_model.log; tail -3 gpurun_out/${R}_beam.err; cut -c1-250 gpurun_out/${R}_beam_K.json gpurun_out/${R}_predict_K.json gpurun_out/${R}_predict_K_simtptr.json
