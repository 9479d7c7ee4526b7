mkdir -p gpurun_out
R=r75
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 300 -k "image_prep" 2>&1 | tail -40) > gpurun_out/${R}_kernels.log
(timeout 900 python bench.py --workload geo_e2e_b256 --steps 20 --warmup 3 2> gpurun_out/${R}_bench_e2e.err | tail -n 1) > gpurun_out/${R}_bench_geo_e2e.json
(timeout 900 python bench.py --steps 30 --warmup 3 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
tail -n 5 gpurun_out/${R}_kernels.log; tail -n 5 gpurun_out/${R}_bench_e2e.err gpurun_out/${R}_bench.err; cut -c1-1800 gpurun_out/${R}_bench_geo_e2e.json; cut -c1-300 gpurun_out/${R}_bench.json
