mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "mha" --timeout 600 2>&1 | tail -30) > gpurun_out/r8_kernels.log
(timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short --timeout 600 2>&1 | tail -40) > gpurun_out/r8_model.log
(timeout 600 python tools/microbench.py 2>&1 | tail -17) > gpurun_out/r8_micro.log
(timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/r8_bench.err | tail -1) > gpurun_out/r8_bench.json
tail -c 2000 gpurun_out/r8_bench.err > gpurun_out/r8_bench.err.tail; rm -f gpurun_out/r8_bench.err
(timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph 2>&1 | tail -1) > gpurun_out/r8_bench_nograph.json
for f in gpurun_out/r8_*; do echo "### $f"; tail -n 20 $f | cut -c1-300; done
