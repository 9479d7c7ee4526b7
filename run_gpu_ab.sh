mkdir -p gpurun_out
R=r25
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -8) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
python tools/step_prof.py 2 > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 320 --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 2 > gpurun_out/${R}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pointer_bwd_mma|add_ln_bwd_fast|add_ln_fwd_fast" -s 20 -c 6 -o gpurun_out/${R}_misc -f python tools/step_prof.py 2 > gpurun_out/${R}_ncu2.log 2>&1
tail -3 gpurun_out/${R}_tests.log; cut -c1-300 gpurun_out/${R}_bench.json; tail -5 gpurun_out/${R}_bench.err; tail -2 gpurun_out/${R}_ncu2.log
