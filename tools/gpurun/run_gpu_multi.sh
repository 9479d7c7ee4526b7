mkdir -p gpurun_out
R=${R:-r02q}
N=${1:-2}
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 2> gpurun_out/${R}_bench${N}.err | tail -n 1) > gpurun_out/${R}_bench${N}.json; echo "default rc=$?"
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2> gpurun_out/${R}_ref${N}.err | tail -n 1) > gpurun_out/${R}_ref${N}.json; echo "reference arm rc=$?"
(timeout 300 python -m pytest tests/test_gpu_baseline_parity.py -m gpu -q --tb=short --timeout 280 -k "two_gpu" 2>&1 | tail -n 6) > gpurun_out/${R}_nccl_test.log
cut -c1-260 gpurun_out/${R}_bench${N}.json; cut -c1-200 gpurun_out/${R}_ref${N}.json; tail -n 3 gpurun_out/${R}_nccl_test.log
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${R}_bench${N}.err | tail -n 4 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/${R}_bench${N}.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['greedy_decode']['value'], d['beam5_decode']['value'], {k:v['value'] for k,v in d['configs'].items()})
PY
