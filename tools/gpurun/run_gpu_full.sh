mkdir -p gpurun_out
R=r23
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -8) > gpurun_out/${R}_tests.log
(timeout 300 python tools/microbench.py 2>&1) > gpurun_out/${R}_micro.log
(timeout 600 python bench.py --steps 20 --warmup 3 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/${R}_ref.err | tail -1) > gpurun_out/${R}_ref.json
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K.json
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 320 --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 2 > gpurun_out/${R}_ncu1.log 2>&1
tail -3 gpurun_out/${R}_tests.log; cut -c1-400 gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_ref.json; cat gpurun_out/${R}_predict_K.json; tail -2 gpurun_out/${R}_ncu1.log
