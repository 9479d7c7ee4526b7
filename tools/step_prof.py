"""One eager (no CUDA graph) train step of the bench workload, for ncu launch lists / captures: python tools/step_prof.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ickb200 import synthetic as syn  # noqa: E402
from ickb200.trainer import Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = syn.BASELINE_CONFIGS[bench.CFG_NAME]
dev = torch.device("cuda", 0)
dec = bench.build_decoder(cfg, dev, torch.bfloat16)
tr = Trainer(dec, lr=4e-4, grad_clip=5.0, use_graph=False)
inp = tr.prepare(*bench.args_of(cfg, bench.host_batch(cfg, seed=0, pin=False)))
for _ in range(steps - 1):
    tr.step(inp)
torch.cuda.synchronize()
# the last step is the profiled one: ncu --nvtx --nvtx-include "profile_step/" captures exactly one whole step
torch.cuda.nvtx.range_push("profile_step")
tr.step(inp)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", float(tr.loss_acc[0] / tr.loss_acc[1]))
