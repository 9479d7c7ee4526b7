mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "mha or pointer" 2>&1 | tail -30) > gpurun_out/r14_kernels.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -10) > gpurun_out/r14_model.log
(timeout 600 python tools/microbench.py 2>&1 | tail -17) > gpurun_out/r14_micro.log
(timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r14_bench.err | tail -1) > gpurun_out/r14_bench.json
tail -c 2000 gpurun_out/r14_bench.err > gpurun_out/r14_bench.err.tail; rm -f gpurun_out/r14_bench.err
for f in gpurun_out/r14_*; do echo "### $f"; tail -n 20 $f | cut -c1-400; done
