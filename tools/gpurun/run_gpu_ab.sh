mkdir -p gpurun_out
R=${R:-r02h}
(timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "few_queries or beam or pointer or grouped or indicators" 2>&1 | tail -n 15) > gpurun_out/${R}_kernels.log
tail -n 6 gpurun_out/${R}_kernels.log
(timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 500 -k "beam or predict or greedy or decode" 2>&1 | tail -n 15) > gpurun_out/${R}_model.log
tail -n 6 gpurun_out/${R}_model.log
for cfg in "" "ICK_ATTN_SPLIT=0" "ICK_PTR_STEP=0"; do
  (env $cfg timeout 200 python tools/bench_predict.py --variant K --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1 | cut -c1-120) >> gpurun_out/${R}_beam_ab.txt
  (env $cfg timeout 200 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1 | cut -c1-120) >> gpurun_out/${R}_greedy_ab.txt
done
tail -n 4 gpurun_out/${R}.err; cat gpurun_out/${R}_beam_ab.txt gpurun_out/${R}_greedy_ab.txt
ICKB200_DECODE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${R}_beam_launches.csv python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 1 gpurun_out/${R}_ncu2.log | cut -c1-100
