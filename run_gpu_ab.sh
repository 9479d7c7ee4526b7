mkdir -p gpurun_out
R=r32
for i in 1 2; do
(ICKB200_LIB=$PWD/alt/libickb200_oldattn.so timeout 300 python tools/microbench.py 2>&1 | tail -16 | grep -E "entity|cross") > gpurun_out/${R}_micro_old$i.log
(timeout 300 python tools/microbench.py 2>&1 | tail -16 | grep -E "entity|cross") > gpurun_out/${R}_micro_new$i.log
done
(ICKB200_LIB=$PWD/alt/libickb200_oldattn.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench_old.err | tail -1) > gpurun_out/${R}_bench_old.json
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
for f in gpurun_out/${R}_micro_*; do echo $f; cat $f; done; cut -c1-200 gpurun_out/${R}_bench_old.json;  cut -c1-200 gpurun_out/${R}_bench.json
