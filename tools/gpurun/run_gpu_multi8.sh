mkdir -p gpurun_out
R=r87
N=${1:-8}
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 3 --workload news_b8 --no-decode --no-trim-extra --no-cpu-baseline 2> gpurun_out/${R}_bench${N}_news.err | tail -n 1) > gpurun_out/${R}_bench${N}_news.json; echo "rc=$?"
cut -c1-300 gpurun_out/${R}_bench${N}_news.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}_news.err | tail -n 4 | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench${N}_news.json').read()); print(d['value'], d['e2e'], d['config']['cuda_graph'], d['config']['global_batch'])"
