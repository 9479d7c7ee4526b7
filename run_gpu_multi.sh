mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r7_gpus.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r7_bench2.err | tail -1) > gpurun_out/r7_bench2.json
tail -c 3000 gpurun_out/r7_bench2.err > gpurun_out/r7_bench2.err.tail; rm -f gpurun_out/r7_bench2.err
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1) > gpurun_out/r7_ref2.json
for f in gpurun_out/r7_*; do echo "### $f"; tail -n 12 $f | cut -c1-500; done
