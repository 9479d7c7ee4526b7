mkdir -p gpurun_out
R=r65
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 -k "predict or decode" 2>&1 | tail -4) > gpurun_out/${R}_tests.log
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K.json
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K2.json
tail -n 3 gpurun_out/${R}_tests.log; cat gpurun_out/${R}_predict_K.json gpurun_out/${R}_predict_K2.json | cut -c1-200
