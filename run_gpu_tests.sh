mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "mha or wgrad" --timeout 600 2>&1 | tail -40) > gpurun_out/r4_kernels.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -40) > gpurun_out/r4_model.log
(timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/r4_bench.err | tail -1) > gpurun_out/r4_bench.json
tail -c 2000 gpurun_out/r4_bench.err > gpurun_out/r4_bench.err.tail; rm -f gpurun_out/r4_bench.err
for f in gpurun_out/r4_*; do echo "### $f"; tail -n 8 $f | cut -c1-400; done
