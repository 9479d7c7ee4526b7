"""
Fused train step for the caption decoder: the reference's train.py step (G/train.py:263-297 — decoder forward, packed
cross-entropy, backward, elementwise gradient clamp at +-5, Adam) without autograd, host syncs or per-parameter launches:

    forward kernels -> fused masked-CE (loss + dscores) -> hand-written backward into ONE flat fp32 gradient buffer
    -> [data-parallel: one NCCL all-reduce of that buffer, which also carries the loss sum and the token count]
    -> one kernel: scale by 1/tokens, clamp, Adam, and re-pack the bf16 / transposed operand copies.

The ~280 launches of a step are captured ONCE into a CUDA graph (``use_graph=True``) and replayed: the step counter
(dropout seed offset, Adam bias corrections) and the learning rate live in device memory, so a replay draws new dropout
masks and applies the right corrections without any host work beyond copying the next batch into the static inputs.

Data-parallel semantics: captions are independent samples, so ranks shard the batch; the reference
clamps the FULL-batch mean gradient, therefore gradients are summed un-normalised, all-reduced BEFORE the clamp, and
divided by the GLOBAL kept-token count inside the optimizer kernel — the single-process large-batch update up to fp32
summation order.
"""
from __future__ import annotations

from types import SimpleNamespace as NS
from typing import Optional

import torch


class Trainer:
    def __init__(self, decoder, lr: float = 4e-4, betas=(0.9, 0.999), eps: float = 1e-8, grad_clip: Optional[float] = 5.0,
                 process_group=None, distributed: bool = False, use_graph: bool = False, trim_padding: bool = False,
                 overlap_allreduce: bool = False):
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
        self.betas, self.eps = betas, eps
        self.clip = float(grad_clip) if grad_clip else 0.0
        self.distributed = distributed
        self.pg = process_group
        self.seed_base = int(torch.initial_seed()) & 0x7FFFFFFF
        self.rank = 0
        if distributed:
            import torch.distributed as dist

            # Replicas must start identical: rank 0's parameters (and moments) win, whatever each rank seeded.  This first
            # collective also creates the NCCL communicator outside of any graph capture.  The dropout seed is rank-dependent:
            # with one seed every rank would draw the SAME masks for its different shard (rows are numbered per rank).
            self.rank = dist.get_rank(process_group)
            for t in (self.eng.P, self.m, self.v):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
            self.eng.repack()
            self.seed_base = (self.seed_base + self.rank * 0x9E3779B1) & 0x7FFFFFFF
        # device-resident step state: read by the kernels at run time (graph replay needs no new kernel arguments)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), lr, dtype=torch.float32, device=dev)
        self._lr = lr
        # Region-wise all-reduce on a side stream while the backward still runs (see _reduce_region).  OFF by default: measured on
        # 2 x B200 it LOSES (5.80 vs 5.72 ms/step): the step's kernels are persistent, one CTA per SM with all of its shared memory,
        # so NCCL's CTAs cannot co-reside - they take SMs at kernel boundaries and the next 148-CTA kernel runs a second wave.
        self.overlap_allreduce = overlap_allreduce
        self._skip_collective = False  # set around the capture warm-up step (see _capture)
        self._regions = None           # engine.grad_regions(), built on first use
        self._comm_stream = None       # side stream of the overlapped gradient all-reduce
        self.use_graph = use_graph
        self._graph = None
        self._static = None
        self._graphs = {}  # captured graphs by caption width (trim_padding produces one width per bucket of 8 positions)
        # Dynamic padding: the reference pads every caption to the dataset maximum (T = max_len + 2, K/create_input_files.py:347)
        # and lets ignore_index drop the <pad> targets.  Positions behind the longest caption of a BATCH feed nothing into the
        # loss and, the decoder being causal, nothing into any gradient, so cutting the batch to that width (rounded up to 8)
        # gives the same loss and the same update.  Off by default: the default step does the reference's full-width work.
        self.trim_padding = trim_padding
        self.pad_id = decoder.word_map.get("<pad>", 0) if hasattr(decoder, "word_map") else 0

    @property
    def lr(self) -> float:
        return self._lr

    def adjust_learning_rate(self, shrink_factor: float) -> None:
        """ut.adjust_learning_rate, G/utils.py:87-97."""
        self._lr *= shrink_factor
        self.lr_dev.fill_(self._lr)

    # ---- checkpoint / resume (the reference pickles the optimizer object next to the modules, G/utils.py:32-46, G/train.py:102-129) ----
    def state_dict(self) -> dict:
        """Everything the fused optimizer needs to resume: Adam moments in the decoder's flat parameter order, the step counter
        (bias corrections, dropout seed offset), learning rate and seed base.  The weights themselves live in the decoder."""
        return {"m": self.m.detach().cpu().clone(), "v": self.v.detach().cpu().clone(), "step": int(self.step_dev.item()),
                "lr": float(self._lr), "betas": tuple(self.betas), "eps": float(self.eps), "clip": float(self.clip),
                "seed_base": int(self.seed_base), "rank": int(self.rank), "n_params": int(self.n)}

    def load_state_dict(self, sd: dict) -> None:
        if int(sd["n_params"]) != self.n:
            raise ValueError(f"optimizer state for {sd['n_params']} parameters does not fit a decoder with {self.n}")
        self.m.copy_(sd["m"].to(self.m.device))
        self.v.copy_(sd["v"].to(self.v.device))
        self.step_dev.fill_(int(sd["step"]))
        self._lr = float(sd["lr"])
        self.lr_dev.fill_(self._lr)
        self.betas, self.eps, self.clip = tuple(sd["betas"]), float(sd["eps"]), float(sd["clip"])
        # a checkpoint written by another rank (rank 0 saves) must not hand its rank-mixed seed to this one
        self.seed_base = (int(sd["seed_base"]) + (self.rank - int(sd.get("rank", 0))) * 0x9E3779B1) & 0x7FFFFFFF
        self.invalidate_graphs()  # betas / eps / clip / seed_base are by-value kernel arguments of the captured steps

    def invalidate_graphs(self) -> None:
        """Drop the captured steps (they bake hyper-parameters, the training flag, the frozen set and buffer addresses)."""
        self._graphs.clear()
        self._graph = self._static = None

    def _frozen(self):
        """parameters that the reference would not update (requires_grad False, e.g. fine_tune_embeddings(False)) keep a zero gradient"""
        return tuple(k for k in self.decoder._param_names if not self.decoder._get(k).requires_grad)

    def trimmed_width(self, captions) -> int:
        """Width that keeps every non-<pad> token of the batch (multiple of 8, at least 8).  A host tensor costs nothing; a
        device tensor costs one small D2H sync."""
        T = captions.shape[1]
        nonpad = (captions != self.pad_id).any(dim=0)
        idx = torch.nonzero(nonpad).flatten()
        last = int(idx[-1]) if idx.numel() else 0
        return min(T, max(8, (last + 1 + 7) // 8 * 8))

    def prepare(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None):
        """Host->device moves and the sort-by-length of DecoderTransformer.forward (G/models.py:330-335); no host sync
        (unless trim_padding has to look at captions that already live on the device)."""
        if self.trim_padding:
            Tw = self.trimmed_width(captions)
            if self.distributed:
                # The width must be ONE decision for all ranks: graphs are keyed by the caption shape, and a rank that meets a
                # new width alone would capture while its peers replay.  (Also: equal work per rank.)
                import torch.distributed as dist

                nccl = dist.get_backend(self.pg) == "nccl"
                w = torch.tensor([Tw], dtype=torch.int32, device=self.eng.device if nccl else "cpu")
                dist.all_reduce(w, op=dist.ReduceOp.MAX, group=self.pg)
                Tw = int(w.item())
            if Tw < captions.shape[1]:
                captions, caption_masks = captions[:, :Tw], caption_masks[:, :Tw]
                caption_lengths = caption_lengths.clamp(max=Tw)
        inp, lengths, _ = self.decoder._sorted_inputs(self.eng.device, captions, encoder_out, caption_masks, caption_lengths, entities, facts)
        inp.decode_len = (lengths - 1).to(torch.int32)
        return inp

    # ---- one step, eager ----------------------------------------------------------------------------------------------------
    def _step_impl(self, inp) -> None:
        eng, K = self.eng, self.eng.K
        self.step_dev.add_(1)
        self.gbuf.zero_()
        K.set_seed_source(self.step_dev)  # effective dropout seed = seed_base + step (read on the device)
        reduce = self.distributed and not self._skip_collective
        try:
            scores, ctx = eng.forward(inp, train=self.decoder.training, seed=self.seed_base)
            _, ds = eng.loss(scores, inp.captions, inp.decode_len, loss_acc=self.loss_acc)
            eng.backward(ctx, ds, self.g, need_encoder_grad=False, on_done=self._reduce_region if (reduce and self.overlap_allreduce) else None)
        finally:
            K.set_seed_source(None)
        if reduce and self.overlap_allreduce:
            self._reduce_join()
        elif reduce:
            import torch.distributed as dist

            dist.all_reduce(self.gbuf, group=self.pg)  # one collective: gradients + [loss_sum, kept_tokens]
        for k in self._frozen():
            eng.param(k, self.g).zero_()
        b1, b2 = self.betas
        K.adam_step(eng.P, self.g, self.m, self.v, self._lr, b1, b2, self.eps, 1.0, 1.0, self.clip, self.loss_acc[1:], 1.0, eng.dstA,
                    eng.dstB, eng.dstC, eng.packT, eng.packF, update=True, step_dev=self.step_dev, lr_dev=self.lr_dev)

    # ---- gradient all-reduce overlapped with the backward pass ------------------------------------------------------------------------
    # The flat buffer is reduced region by region, in the order the backward completes them (engine.grad_regions): the score heads'
    # weights (+ the loss sum / token count behind them) while the decoder stack is still being differentiated, the decoder layers
    # and the word embedding under the encoder stacks, each encoder layer under the next one; only the last encoder layer and the
    # small embeddings are exposed.  On CUDA the collectives run on a side stream forked / joined with events (inside a captured
    # step they become a parallel branch of the graph); sums are independent per element, so the result equals one big all-reduce.
    def _reduce_region(self, region: str) -> None:
        import torch.distributed as dist

        if self._regions is None:
            self._regions = self.eng.grad_regions()
        slices = list(self._regions[region])
        if region == "heads":  # the tail of the buffer carries [loss_sum, kept_tokens] right behind the last parameter
            slices[-1] = (slices[-1][0], self.n + 2)
        if self.gbuf.device.type != "cuda":
            for lo, hi in slices:
                dist.all_reduce(self.gbuf[lo:hi], group=self.pg)
            return
        main = torch.cuda.current_stream(self.gbuf.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(self.gbuf.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._comm_stream.wait_event(ev)
        with torch.cuda.stream(self._comm_stream):
            for lo, hi in slices:
                dist.all_reduce(self.gbuf[lo:hi], group=self.pg)

    def _reduce_join(self) -> None:
        if self.gbuf.device.type == "cuda" and self._comm_stream is not None:
            torch.cuda.current_stream(self.gbuf.device).wait_stream(self._comm_stream)

    def step(self, inp) -> torch.Tensor:
        """One optimisation step on a prepared batch.  Returns a device tensor [loss_sum, kept_tokens] (global under DDP)."""
        self.eng = self.decoder._ensure_engine()  # re-gathers / re-packs if the parameter storages were replaced or invalidated
        if not self.use_graph:
            self._step_impl(inp)
            return self.loss_acc
        # everything a captured step bakes in: shapes, the training flag (dropout on / off), the frozen set, and the parameter
        # storage generation (a re-gathered flat buffer moves the fp32 addresses the kernels read)
        key = (tuple(inp.captions.shape), bool(self.decoder.training), self._frozen(), getattr(self.decoder, "_flat_gen", 0))
        if key not in self._graphs:
            self._capture(inp)
            self._graphs[key] = (self._graph, self._static)
        else:
            self._graph, self._static = self._graphs[key]
            for k, v in vars(self._static).items():
                if torch.is_tensor(v):
                    v.copy_(getattr(inp, k), non_blocking=True)
        self._graph.replay()
        return self.loss_acc

    def _capture(self, inp) -> None:
        self._static = NS(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in vars(inp).items()})
        # One eager warm-up step on a side stream (lazy one-time initialisation must not happen during capture); the
        # optimizer state is snapshotted and restored around it so that the first call still performs exactly one step.
        snap = [t.clone() for t in (self.eng.P, self.m, self.v, self.step_dev)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            # no collective in the warm-up: a rank that captures a new shape alone (e.g. a ragged last batch) must still issue
            # exactly ONE all-reduce for this step, like its peers that replay - the warm-up's result is thrown away anyway
            self._skip_collective = True
            try:
                self._step_impl(self._static)
            finally:
                self._skip_collective = False
        torch.cuda.current_stream().wait_stream(s)
        for t, c in zip((self.eng.P, self.m, self.v, self.step_dev), snap):
            t.copy_(c)
        self.eng.repack()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_impl(self._static)  # capture only: nothing executes here
        self._graph = g

    def train_step(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None) -> torch.Tensor:
        return self.step(self.prepare(captions, encoder_out, caption_masks, caption_lengths, entities, facts))

    def validate_step(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None) -> torch.Tensor:
        """validate() of train.py (G/train.py:317-386): eval-mode forward (no dropout) and the packed cross-entropy, no gradient and
        no update.  Returns a fresh device tensor [loss_sum, kept_tokens] (summed over the ranks under DDP); the reference's
        `losses.avg` over an epoch is sum(loss_sum) / sum(kept_tokens) weighted the same way (mean over non-<pad> targets)."""
        inp = self.prepare(captions, encoder_out, caption_masks, caption_lengths, entities, facts)
        eng = self.eng
        with torch.no_grad():
            scores, _ = eng.forward(inp, train=False, seed=None)
            acc, _ = eng.loss(scores, inp.captions, inp.decode_len, want_grad=False)
        if self.distributed:
            import torch.distributed as dist

            dist.all_reduce(acc, group=self.pg)
        return acc

    def run(self, host_batches):
        """
        Steady-state training loop over HOST batches (tuples in train.py's argument order, G/train.py:263-272; pinned memory
        makes the copies asynchronous).  The reference moves a batch to the device and only then starts computing
        (G/train.py:263-269); here the host->device copies of batch i+1 run on a copy stream while step i computes, and the
        [loss_sum, kept_tokens] pair of step i is read back into pinned memory while step i+1 runs.  Yields one host tensor
        of 2 floats per step, in order; every step's inputs are copied and every step's result is read.
        """
        dev = self.eng.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._host_acc = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
        cs = self._copy_stream

        def stage(hb):
            with torch.cuda.stream(cs):
                d = tuple(t.to(dev, non_blocking=True) if torch.is_tensor(t) else t for t in hb)
                ev = torch.cuda.Event()
                ev.record(cs)
            return d, ev

        it = iter(host_batches)
        first = next(it, None)
        nxt = stage(first) if first is not None else None
        pending = None
        i = 0
        while nxt is not None:
            d, ev = nxt
            main.wait_event(ev)
            for t in d:
                if torch.is_tensor(t):
                    t.record_stream(main)  # allocated on the copy stream, consumed on the compute stream
            hb = next(it, None)
            nxt = stage(hb) if hb is not None else None  # overlaps the step launched below
            acc = self.train_step(*d)
            host = self._host_acc[i & 1]
            host.copy_(acc, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].clone()
            pending = (host, done)
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].clone()


def generate_sharded(decoder, encoder_out, max_pred_len, entities, facts=None, beam_size: int = 0, process_group=None):
    """
    Caption generation over several GPUs: images are independent, so rank r decodes the images i with
    i % world_size == r - no communication on the data path - and the token ids are gathered on every rank at the end
    (one all_gather of small int64 tensors) and put back in the original order.  beam_size = 0: greedy predict() semantics
    (DecoderTransformer.predict_batch); > 0: beam search (extension).  Every rank passes the FULL inputs (or at least its shard's
    rows at their global positions); returns (N, max_pred_len) int64 on the CPU.  Without an initialised process group this is the
    single-process call.
    """
    import torch.distributed as dist

    n = encoder_out.shape[0]
    world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(process_group) if world > 1 else 0
    mine = torch.arange(rank, n, world)
    args = (encoder_out[mine.to(encoder_out.device)], max_pred_len, entities[mine.to(entities.device)],
            facts[mine.to(facts.device)] if facts is not None else None)
    if mine.numel() == 0:
        out = torch.zeros((0, max_pred_len), dtype=torch.int64)
    elif beam_size > 0:
        out = decoder.beam_search_batch(*args, beam_size=beam_size).cpu()
    else:
        out = decoder.predict_batch(*args).cpu()
    if world == 1:
        return out
    per_rank = (n + world - 1) // world
    padded = torch.zeros((per_rank, max_pred_len), dtype=torch.int64)
    padded[: out.shape[0]] = out
    if dist.get_backend(process_group) == "nccl":
        dev = decoder._ensure_engine().device  # the inputs may live on the host (predict_batch moves them itself)
        bufs = [torch.empty_like(padded, device=dev) for _ in range(world)]
        dist.all_gather(bufs, padded.to(dev), group=process_group)
        bufs = [b.cpu() for b in bufs]
    else:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(bufs, padded, group=process_group)
    full = torch.zeros((n, max_pred_len), dtype=torch.int64)
    for r in range(world):
        idx = torch.arange(r, n, world)
        full[idx] = bufs[r][: idx.numel()]
    return full
