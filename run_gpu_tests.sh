mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "wgrad" --timeout 300 2>&1 | tail -15) > gpurun_out/r2_wgrad.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -60) > gpurun_out/r2_model_tc.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4) > gpurun_out/r2_smoke.log
(timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2_bench.err | tail -2) > gpurun_out/r2_bench.json
tail -c 3000 gpurun_out/r2_bench.err > gpurun_out/r2_bench.err.tail; rm -f gpurun_out/r2_bench.err
(timeout 600 python bench.py --steps 3 --warmup 3 --dtype fp32 --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/r2_bench_fp32.json
for f in gpurun_out/r2_*; do echo "### $f"; tail -n 6 $f | cut -c1-600; done
