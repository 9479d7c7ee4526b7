mkdir -p gpurun_out
R=r68
(timeout 900 python -m pytest tests -m gpu -q --tb=short -x --timeout 600 2>&1 | tail -4) > gpurun_out/${R}_tests.log
(timeout 300 python tools/bench_predict.py --variant K 2>&1 | tail -1) > gpurun_out/${R}_predict_K.json
(timeout 600 python bench.py --workload news_b8 --steps 50 --warmup 3 --no-cpu-baseline --no-decode --no-trim-extra 2> gpurun_out/${R}_bench_news.err | tail -1) > gpurun_out/${R}_bench_news.json
(timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-decode --no-trim-extra 2> gpurun_out/${R}_bench.err | tail -1) > gpurun_out/${R}_bench.json
tail -n 3 gpurun_out/${R}_tests.log; cut -c1-200 gpurun_out/${R}_predict_K.json; cut -c1-170 gpurun_out/${R}_bench_news.json; cut -c1-170 gpurun_out/${R}_bench.json
