"""Worker of tests/test_gpu_baseline_parity.py::test_two_gpu_nccl_equals_single_gpu: launched under torchrun with 2 ranks, each takes
half of the batch, runs two fused train steps per mode over NCCL and rank 0 saves the resulting parameters."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from helpers import batch_args, build_module  # noqa: E402
from ickb200 import synthetic as syn  # noqa: E402
from ickb200.trainer import Trainer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    cfg = syn.SMALL_CONFIGS["K"].with_batch(4)
    full = syn.make_batch(cfg, seed=4, equal_lengths=False)
    n = cfg.B // world
    shard = {k: v[rank * n:(rank + 1) * n] for k, v in full.items()}
    scfg = cfg.with_batch(n)
    out = {}
    for mode, kw in (("eager", {}), ("graph", {"use_graph": True}), ("overlap", {"overlap_allreduce": True}),
                     ("graph+overlap", {"use_graph": True, "overlap_allreduce": True})):
        dec = build_module(scfg, dev, torch.float32, dropouts=(0.0, 0.0, 0.0), profile="test").train()
        tr = Trainer(dec, lr=4e-4, grad_clip=5.0, distributed=True, **kw)
        accs = [tr.train_step(*batch_args(scfg, shard)).cpu().tolist() for _ in range(2)]
        out[mode] = ({k: v.detach().cpu().clone() for k, v in dec.named_parameters()}, accs)
        tr.invalidate_graphs()
        torch.cuda.synchronize()
    if rank == 0:
        torch.save(out, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
