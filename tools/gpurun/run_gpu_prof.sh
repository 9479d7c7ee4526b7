mkdir -p gpurun_out
R=r93
ICKB200_DECODE_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 200 --csv --log-file gpurun_out/${R}_greedy_launches.csv python tools/bench_predict.py --variant K --reps 1 > gpurun_out/${R}_ncu1.log 2>&1
ICKB200_DECODE_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${R}_beam_launches.csv python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 1 gpurun_out/${R}_ncu1.log | cut -c1-100; tail -n 1 gpurun_out/${R}_ncu2.log | cut -c1-100
