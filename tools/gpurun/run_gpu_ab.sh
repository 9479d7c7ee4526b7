mkdir -p gpurun_out
R=${R:-r02e}
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 200 -k "gemm_add_ln or rowdot or gemm_dual or gemm_plain" 2>&1 | tail -n 25) > gpurun_out/${R}_kernels.log
tail -n 25 gpurun_out/${R}_kernels.log
(timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 600 --deselect tests/test_gpu_kernels.py 2>&1 | tail -n 25) > gpurun_out/${R}_tests.log
tail -n 8 gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 30 --warmup 3 2> gpurun_out/${R}_bench.err | tail -n 1) > gpurun_out/${R}_bench.json
tail -n 5 gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench.json
(ICK_FUSE_LN=0 timeout 600 python bench.py --steps 30 --warmup 3 2> gpurun_out/${R}_bench_nofuse.err | tail -n 1) > gpurun_out/${R}_bench_nofuse.json
cut -c1-300 gpurun_out/${R}_bench_nofuse.json
