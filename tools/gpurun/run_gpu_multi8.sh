mkdir -p gpurun_out
R=r62
N=${1:-8}
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 3 --graph --no-decode --no-trim-extra 2> gpurun_out/${R}_bench${N}g.err | tail -1) > gpurun_out/${R}_bench${N}g.json; echo "rc=$?"
cut -c1-300 gpurun_out/${R}_bench${N}g.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}g.err | tail -n 5 | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench${N}g.json').read()); print(d['e2e'], d['config']['cuda_graph'])"
