mkdir -p gpurun_out
R=r48
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x --timeout 300 -k "mha" 2>&1 | tail -4) > gpurun_out/${R}_kernels.log
(ICKB200_LIB=$PWD/alt/libickb200_before.so timeout 300 python tools/microbench.py 2>&1 | tail -16 | grep -E "bwd") > gpurun_out/${R}_micro_old.log
(timeout 300 python tools/microbench.py 2>&1 | tail -16 | grep -E "bwd") > gpurun_out/${R}_micro_new.log
for i in 1 2; do
(ICKB200_LIB=$PWD/alt/libickb200_before.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode 2> gpurun_out/${R}_bench_old.err | tail -1) > gpurun_out/${R}_bench_old$i.json
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench$i.json
done
cat gpurun_out/${R}_kernels.log; for f in gpurun_out/${R}_micro_*; do echo $f; cat $f; done; for i in 1 2; do cut -c1-160 gpurun_out/${R}_bench_old$i.json;  cut -c1-160 gpurun_out/${R}_bench$i.json; done; tail -n 3 gpurun_out/${R}_bench.err
