mkdir -p gpurun_out
R=r85
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -k "indicators or grouped" 2>&1 | tail -n 10) > gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -n 10) > gpurun_out/${R}_model.log
for v in N K; do
(timeout 600 python tools/bench_predict.py --variant $v --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_$v.json
(timeout 600 python tools/bench_predict.py --variant $v --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_beam_$v.json
done
(timeout 600 python bench.py --workload news_b8 --steps 30 --no-cpu-baseline --no-decode 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_bench_news_b8.json
tail -n 3 gpurun_out/${R}_kernels.log gpurun_out/${R}_model.log gpurun_out/${R}.err; cut -c1-250 gpurun_out/${R}_*.json
