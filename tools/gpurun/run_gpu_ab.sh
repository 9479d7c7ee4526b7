mkdir -p gpurun_out
R=r60
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K.json
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K2.json
cat gpurun_out/${R}_predict_K.json gpurun_out/${R}_predict_K2.json | cut -c1-200
