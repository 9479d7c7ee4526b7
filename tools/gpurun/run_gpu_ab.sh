mkdir -p gpurun_out
R=r78
D=$PWD/image-captioning-with-external-knowledge_b200/csrc
python - <<'PY' > gpurun_out/r78_readbw.log 2>&1
import torch
x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda").normal_()
for name, fn in (("sum", lambda: x.sum()), ("amax", lambda: x.amax())):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(name, "read-only GB/s", x.numel() * 2 / best / 1e6)
PY
cat gpurun_out/r78_readbw.log
for v in "" _kg4u2 _kg4u4 _kg2u4; do
  (ICKB200_LIB=$D/libickb200$v.so timeout 600 python tools/bench_predict.py --variant K --reps 5 2>> gpurun_out/${R}.err | tail -n 1) > gpurun_out/${R}_predict_K$v.json
done
tail -n 3 gpurun_out/${R}.err; cut -c1-230 gpurun_out/${R}_predict*.json
