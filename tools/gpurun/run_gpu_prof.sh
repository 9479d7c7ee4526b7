mkdir -p gpurun_out
R=${R:-r02f}
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 200 -k "ce_adam or gemm_add_ln or rowdot" 2>&1 | tail -n 6) > gpurun_out/${R}_kernels.log
tail -n 4 gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 600 --deselect tests/test_gpu_kernels.py 2>&1 | tail -n 25) > gpurun_out/${R}_tests.log
tail -n 6 gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 30 --warmup 3 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
tail -n 5 gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench.json
ICKB200_DECODE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 200 --csv --log-file gpurun_out/${R}_greedy_launches.csv python tools/bench_predict.py --variant K --reps 1 > gpurun_out/${R}_ncu1.log 2>&1
ICKB200_DECODE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/${R}_beam_launches.csv python tools/bench_predict.py --variant K --beam 5 --reps 1 > gpurun_out/${R}_ncu2.log 2>&1
tail -n 1 gpurun_out/${R}_ncu1.log | cut -c1-100; tail -n 1 gpurun_out/${R}_ncu2.log | cut -c1-100
