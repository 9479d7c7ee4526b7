mkdir -p gpurun_out
R=r37
for i in 1 2; do
(ICK_ATTN_DS=0 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench_old.err | tail -1) > gpurun_out/${R}_bench_old$i.json
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench$i.json
done
for i in 1 2; do cut -c1-200 gpurun_out/${R}_bench_old$i.json;  cut -c1-200 gpurun_out/${R}_bench$i.json; done; tail -n 3 gpurun_out/${R}_bench.err
