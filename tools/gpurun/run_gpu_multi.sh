mkdir -p gpurun_out
R=r81
N=${1:-2}
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2> gpurun_out/${R}_bench${N}.err | tail -n 1) > gpurun_out/${R}_bench${N}.json; echo "default rc=$?"
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --workload geo_e2e_b256 2> gpurun_out/${R}_bench${N}_geo_e2e.err | tail -n 1) > gpurun_out/${R}_bench${N}_geo_e2e.json; echo "geo_e2e rc=$?"
for f in gpurun_out/${R}_bench${N}.json gpurun_out/${R}_bench${N}_geo_e2e.json; do cut -c1-260 $f; done
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}.err | tail -n 4 | cut -c1-300; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}_geo_e2e.err | tail -n 4 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${R}_bench${N}.json').read()); print(d['value'], d['e2e']['value'], d['greedy_decode']['value'], d['beam5_decode']['value'])
d=json.loads(open('gpurun_out/${R}_bench${N}_geo_e2e.json').read()); print(d['value'], d['e2e']['value'], d['breakdown_ms'])
PY
