mkdir -p gpurun_out
R=r66
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 -k "predict or decode" 2>&1 | tail -4) > gpurun_out/${R}_tests.log
(ICK_DECODE_ROWS=0 timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K_old.json
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K.json
(timeout 300 python tools/bench_predict.py --variant N 2>&1 | tail -1) > gpurun_out/${R}_predict_N.json
tail -n 3 gpurun_out/${R}_tests.log; cat gpurun_out/${R}_predict_K_old.json gpurun_out/${R}_predict_K.json gpurun_out/${R}_predict_N.json | cut -c1-330
