"""
Fused train step for the caption decoder: the reference's train.py step (G/train.py:263-297 — decoder forward, packed
cross-entropy, backward, elementwise gradient clamp at +-5, Adam) without autograd, host syncs or per-parameter launches:

    forward kernels -> fused masked-CE (loss + dscores) -> hand-written backward into ONE flat fp32 gradient buffer
    -> [data-parallel: one NCCL all-reduce of that buffer, which also carries the loss sum and the token count]
    -> one kernel: scale by 1/tokens, clamp, Adam, and re-pack the bf16 / transposed operand copies.

Data-parallel semantics: captions are independent samples, so ranks shard the batch; the reference
clamps the FULL-batch mean gradient, therefore gradients are summed un-normalised, all-reduced BEFORE the clamp, and
divided by the GLOBAL kept-token count inside the optimizer kernel — bit-for-bit the single-process large-batch update
up to fp32 summation order.
"""
from __future__ import annotations

from types import SimpleNamespace as NS
from typing import Optional

import torch


class Trainer:
    def __init__(self, decoder, lr: float = 4e-4, betas=(0.9, 0.999), eps: float = 1e-8, grad_clip: Optional[float] = 5.0,
                 process_group=None, distributed: bool = False):
        self.decoder = decoder
        self.eng = decoder._ensure_engine()
        n = self.eng.plan.n_params
        dev = self.eng.device
        self.n = n
        # gradients + [loss_sum, kept_tokens] in one buffer so that one all-reduce moves everything
        self.gbuf = torch.zeros(n + 2, dtype=torch.float32, device=dev)
        self.g = self.gbuf[:n]
        self.loss_acc = self.gbuf[n:]
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.clip = float(grad_clip) if grad_clip else 0.0
        self.t = 0
        self.distributed = distributed
        self.pg = process_group
        self.seed_base = int(torch.initial_seed()) & 0x7FFFFFFF
        # parameters that the reference would not update (requires_grad False) keep a zero gradient
        self._frozen = [k for k in decoder._param_names if not decoder._get(k).requires_grad]

    def prepare(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None):
        """Host->device moves and the sort-by-length of DecoderTransformer.forward (G/models.py:330-335); no host sync."""
        inp, lengths, _ = self.decoder._sorted_inputs(self.eng.device, captions, encoder_out, caption_masks, caption_lengths, entities, facts)
        inp.decode_len = (lengths - 1).to(torch.int32)
        return inp

    def step(self, inp) -> torch.Tensor:
        """One optimisation step on a prepared batch.  Returns a device tensor [loss_sum, kept_tokens] (global under DDP)."""
        eng, K = self.eng, self.eng.K
        self.t += 1
        seed = (self.seed_base * 1000003 + self.t) & 0x7FFFFFFF
        self.gbuf.zero_()
        scores, ctx = eng.forward(inp, train=self.decoder.training, seed=seed)
        _, ds = eng.loss(scores, inp.captions, inp.decode_len, loss_acc=self.loss_acc)
        eng.backward(ctx, ds, self.g, need_encoder_grad=False)
        for k in self._frozen:
            eng.param(k, self.g).zero_()
        if self.distributed:
            import torch.distributed as dist

            dist.all_reduce(self.gbuf, group=self.pg)
        b1, b2 = self.betas
        K.adam_step(eng.P, self.g, self.m, self.v, self.lr, b1, b2, self.eps, 1.0 - b1 ** self.t, 1.0 - b2 ** self.t, self.clip,
                    self.loss_acc[1:], 1.0, eng.dstA, eng.dstB, eng.dstC, eng.packT, eng.packF, update=True)
        return self.loss_acc

    def train_step(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None) -> torch.Tensor:
        return self.step(self.prepare(captions, encoder_out, caption_masks, caption_lengths, entities, facts))

    def adjust_learning_rate(self, shrink_factor: float) -> None:
        """ut.adjust_learning_rate, G/utils.py:87-97."""
        self.lr *= shrink_factor
