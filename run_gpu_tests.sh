mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "gemm or wgrad" --timeout 600 2>&1 | tail -40) > gpurun_out/r6_kernels.log
(timeout 600 python tools/microbench.py 2>&1 | head -26) > gpurun_out/r6_micro.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -20) > gpurun_out/r6_model.log
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r6_bench.err | tail -1) > gpurun_out/r6_bench.json
tail -c 2000 gpurun_out/r6_bench.err > gpurun_out/r6_bench.err.tail; rm -f gpurun_out/r6_bench.err
for f in gpurun_out/r6_*; do echo "### $f"; tail -n 30 $f | cut -c1-300; done
