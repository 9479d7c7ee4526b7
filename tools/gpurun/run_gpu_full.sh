mkdir -p gpurun_out
R=r80
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --timeout 900 2>&1 | tail -n 15) > gpurun_out/${R}_tests.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 2) > gpurun_out/${R}_smoke.log
(timeout 900 python bench.py 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
ICKB200_DECODE_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:mha_decode_tma -s 7 -c 1 -o gpurun_out/${R}_decode_tma python tools/bench_predict.py --variant K --reps 1 > gpurun_out/${R}_ncu1.log 2>&1
tail -n 4 gpurun_out/${R}_tests.log gpurun_out/${R}_smoke.log gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench.json; tail -n 2 gpurun_out/${R}_ncu1.log
