mkdir -p gpurun_out
R=r30
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -4) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/${R}_smoke.log
tail -n 3 gpurun_out/${R}_tests.log; cut -c1-900 gpurun_out/${R}_bench.json; tail -n 5 gpurun_out/${R}_bench.err; cat gpurun_out/${R}_smoke.log
