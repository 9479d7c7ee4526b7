mkdir -p gpurun_out
R=r56
(timeout 300 python tools/microbench.py 2>&1 | head -14) > gpurun_out/${R}_micro_old.log
(ICK_GEMM_DIRECT_SMALL=1 timeout 300 python tools/microbench.py 2>&1 | head -14) > gpurun_out/${R}_micro_new.log
for i in 1 2; do
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode --no-trim-extra 2> gpurun_out/${R}_bench_old.err | tail -1) > gpurun_out/${R}_bench_old$i.json
(ICK_GEMM_DIRECT_SMALL=1 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode --no-trim-extra 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench$i.json
done
paste gpurun_out/${R}_micro_old.log gpurun_out/${R}_micro_new.log | cut -c1-240; for i in 1 2; do cut -c1-160 gpurun_out/${R}_bench_old$i.json;  cut -c1-160 gpurun_out/${R}_bench$i.json; done
