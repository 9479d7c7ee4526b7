mkdir -p gpurun_out
R=r86
for n in 1 2 3 4; do
(ICKB200_DECODE_STREAMS=$n timeout 600 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_K_s$n.json
(ICKB200_DECODE_STREAMS=$n timeout 600 python tools/bench_predict.py --variant K --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_beam_K_s$n.json
done
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 -k "beam or predict" 2>&1 | tail -n 6) > gpurun_out/${R}_model.log
tail -n 3 gpurun_out/${R}.err gpurun_out/${R}_model.log; cut -c1-330 gpurun_out/${R}_*_s*.json
