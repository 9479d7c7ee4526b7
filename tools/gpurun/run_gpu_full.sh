mkdir -p gpurun_out
R=${R:-r02i}
(timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 2>&1 | tail -n 40) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 50 --warmup 3 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench_reference_arm.json
(timeout 200 python tools/step_prof.py 3 2>&1 | tail -n 1) > gpurun_out/${R}_step_plain.log
timeout 600 ncu --nvtx --nvtx-include "profile_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python tools/step_prof.py 3 > gpurun_out/${R}_ncu.log 2>&1
tail -n 12 gpurun_out/${R}_tests.log; tail -n 5 gpurun_out/${R}_bench.err; cut -c1-400 gpurun_out/${R}_bench.json; cat gpurun_out/${R}_step_plain.log; tail -n 2 gpurun_out/${R}_ncu.log
