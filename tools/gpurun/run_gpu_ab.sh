mkdir -p gpurun_out
R=r54
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -5) > gpurun_out/${R}_tests.log
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
tail -n 5 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_bench.json; tail -n 3 gpurun_out/${R}_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/${R}_bench.json').read()); print(d['trimmed_padding']); print(d['loss'], d['kept_tokens'])"
