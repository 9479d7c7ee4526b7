mkdir -p gpurun_out
R=r63
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -3) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/${R}_ref.err | tail -1) > gpurun_out/${R}_ref.json
tail -n 3 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_bench.json; cut -c1-160 gpurun_out/${R}_ref.json; tail -n 3 gpurun_out/${R}_bench.err
