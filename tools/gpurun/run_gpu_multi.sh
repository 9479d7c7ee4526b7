mkdir -p gpurun_out
R=r39
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 2> gpurun_out/${R}_bench2.err | tail -1) > gpurun_out/${R}_bench2.json; echo "eager rc=$?"
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 3 --graph 2> gpurun_out/${R}_bench2g.err | tail -1) > gpurun_out/${R}_bench2g.json; echo "graph rc=$?"
for f in gpurun_out/${R}_bench2.json gpurun_out/${R}_bench2g.json; do cut -c1-330 $f; done
tail -n 5 gpurun_out/${R}_bench2.err | cut -c1-300; tail -n 5 gpurun_out/${R}_bench2g.err | cut -c1-300
