mkdir -p gpurun_out
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r9_bench2.err | tail -1) > gpurun_out/r9_bench2.json
tail -c 1500 gpurun_out/r9_bench2.err > gpurun_out/r9_bench2.err.tail; rm -f gpurun_out/r9_bench2.err
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --graph 2> gpurun_out/r9_bench2g.err | tail -1) > gpurun_out/r9_bench2g.json
tail -c 1500 gpurun_out/r9_bench2g.err > gpurun_out/r9_bench2g.err.tail; rm -f gpurun_out/r9_bench2g.err
for f in gpurun_out/r9_*; do echo "### $f"; tail -n 12 $f | cut -c1-700; done
