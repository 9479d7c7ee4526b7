mkdir -p gpurun_out
R=r64
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 -k "pointer or golden or bf16 or dropout" 2>&1 | tail -4) > gpurun_out/${R}_tests.log
for i in 1 2; do
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode --no-trim-extra 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench$i.json
done
tail -n 4 gpurun_out/${R}_tests.log; for i in 1 2; do cut -c1-160 gpurun_out/${R}_bench$i.json; done; python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench1.json').read()); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items() if 'pointer' in k})"
