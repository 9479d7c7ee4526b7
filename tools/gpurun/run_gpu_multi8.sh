mkdir -p gpurun_out
R=${R:-r03h}
N=8
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2> gpurun_out/${R}_bench${N}.err | tail -n 1) > gpurun_out/${R}_bench${N}.json; echo "default rc=$?"
cut -c1-260 gpurun_out/${R}_bench${N}.json
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}.err | tail -n 4 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${R}_bench${N}.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['greedy_decode']['value'], d['beam5_decode']['value'], {k:v['value'] for k,v in d['configs'].items()})
PY
