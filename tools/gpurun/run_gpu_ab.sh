mkdir -p gpurun_out
R=${R:-r02n}
(timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "beam" 2>&1 | tail -n 15) > gpurun_out/${R}_kernels.log
tail -n 8 gpurun_out/${R}_kernels.log
(timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 500 -k "beam" 2>&1 | tail -n 15) > gpurun_out/${R}_model.log
tail -n 6 gpurun_out/${R}_model.log
(timeout 200 python tools/bench_predict.py --variant K --beam 5 --reps 5 2>> gpurun_out/${R}.err | tail -n 1 | cut -c1-120) >> gpurun_out/${R}_beam_ab.txt
tail -n 4 gpurun_out/${R}.err; cat gpurun_out/${R}_beam_ab.txt
ICKB200_DECODE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${R}_beam_launches.csv python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 1 gpurun_out/${R}_ncu2.log | cut -c1-100
