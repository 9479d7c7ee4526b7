mkdir -p gpurun_out
R=r83
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -k "decode_chain" 2>&1 | tail -n 30) > gpurun_out/${R}_kernels.log
for m in 1 0; do
(ICK_DECODE_CHAIN=$m timeout 600 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_K_chain$m.json
done
ICKB200_DECODE_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/${R}_decode_launches.csv python tools/bench_predict.py --variant K --reps 1 > gpurun_out/${R}_ncu1.log 2>&1
tail -n 5 gpurun_out/${R}_kernels.log; tail -n 3 gpurun_out/${R}.err; cut -c1-200 gpurun_out/${R}_*chain*.json
