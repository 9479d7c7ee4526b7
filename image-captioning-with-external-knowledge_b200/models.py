"""
Drop-in module API: ``Encoder`` and ``DecoderTransformer`` with the reference's constructor and call signatures
(G/models.py:9-60, 212-443; K/models.py:290-609; N/models.py:273-592) and the reference's ``state_dict`` key names,
backed by the sm_100a kernel engine.  Variant modules: ``ickb200.geo_aware``, ``ickb200.knowledge_aware``,
``ickb200.news_knowledge_aware`` (each exports ``Encoder``, ``DecoderTransformer``, ``device``).

    decoder(captions, encoder_out, caption_masks, caption_lengths, entities[, facts])
        -> (scores (B,T,V+E[+F]) fp32, captions_sorted (B,T) int64, decode_lengths list[int])     G/train.py:270-272
    decoder.predict(encoder_out, max_pred_len, entities[, facts]) -> (max_pred_len, 1) int64     G/eval.py:83

There is no CPU path: calling the decoder without a CUDA device (or without the built kernel library) raises.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace as NS
from typing import List, Optional

import torch
from torch import nn

from . import layout
from .engine import DecoderEngine

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")  # same module-level name as the reference


def _default_dtype() -> torch.dtype:
    return {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}[
        os.environ.get("ICKB200_DTYPE", "bf16").lower()]


class Encoder(nn.Module):
    """
    Image encoder (G/models.py:9-60): ResNet-101 trunk -> AdaptiveAvgPool2d(14) -> 1x1 conv 2048->emb_dim -> (B, emb_dim, 196).
    The trunk is stock torchvision/cuDNN (out of scope of the B200 path).
    """

    def __init__(self, encoded_image_size=14, emb_dim=300, encoder_dim=2048, pretrained=True, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        import torchvision

        self.emb_dim = emb_dim
        self.compute_dtype = compute_dtype  # dtype of the hand-off GEMM (None: ICKB200_DTYPE / bf16)
        if pretrained:  # the reference always asks for the ImageNet weights (G/models.py:24)
            try:
                resnet = torchvision.models.resnet101(weights=torchvision.models.ResNet101_Weights.IMAGENET1K_V1)
            except Exception as e:  # no network and no cached weights
                if os.environ.get("ICKB200_ALLOW_RANDOM_TRUNK") != "1":
                    raise RuntimeError(
                        "Encoder(pretrained=True): the ImageNet ResNet-101 weights could not be loaded "
                        f"({type(e).__name__}: {e}).  The reference always starts from them (G/models.py:24); pass "
                        "pretrained=False (or set ICKB200_ALLOW_RANDOM_TRUNK=1) to build a randomly initialised trunk on purpose."
                    ) from e
                import warnings

                warnings.warn(f"Encoder: ImageNet weights unavailable ({e}); RANDOM trunk (ICKB200_ALLOW_RANDOM_TRUNK=1)")
                resnet = torchvision.models.resnet101(weights=None)
        else:
            resnet = torchvision.models.resnet101(weights=None)
        self.resnet = nn.Sequential(*list(resnet.children())[:-2])
        self.adaptive_pool = nn.AdaptiveAvgPool2d((encoded_image_size, encoded_image_size))
        self.conv1 = nn.Conv2d(encoder_dim, emb_dim, 1)
        self.fine_tune()

    def forward(self, images):
        return self.head(self.resnet(images))

    def head(self, feats):
        """
        Trunk output (B, encoder_dim, h, w) -> (B, emb_dim, 14*14): AdaptiveAvgPool2d + 1x1 conv + view (G/models.py:43-46).
        On a CUDA device this is ALWAYS the B200 hand-off path: one pooling kernel that writes the GEMM's K-major A operand, the
        1x1 convolution as a tcgen05 GEMM with the bias in its epilogue, one transpose to the channel-major layout the reference
        returns.  Under autograd - the reference's own call sites run with gradients enabled and conv1.requires_grad = True
        (G/train.py:269, G/eval.py:77-83, G/models.py:32) - the same kernels run inside one autograd node (_EncoderHeadFn) whose
        backward is hand-written: conv1's weight / bias gradient on the tensor-core wgrad kernel, and, when the trunk is being
        fine-tuned, the input gradient (dgrad GEMM + pooling backward).  A CPU tensor takes the stock ops (the trunk itself is
        stock PyTorch and may legitimately run anywhere; the decoder has no CPU path).
        """
        if not feats.is_cuda:
            out = self.conv1(self.adaptive_pool(feats))
            return out.view(out.shape[0], self.emb_dim, -1)
        needs_grad = torch.is_grad_enabled() and (feats.requires_grad or self.conv1.weight.requires_grad or self.conv1.bias.requires_grad)
        if needs_grad:
            return _EncoderHeadFn.apply(self, feats, self.conv1.weight, self.conv1.bias)
        return self._head_fwd(feats)[0]

    def _kernels_and_dtype(self):
        from .kernels import CudaKernels

        K = self.__dict__.get("_kernels")
        if K is None:
            K = self.__dict__["_kernels"] = CudaKernels()
        return K, (getattr(self, "compute_dtype", None) or _default_dtype())

    def repack(self) -> None:
        """Drop the packed copy of conv1 (see DecoderTransformer.repack: needed after writes through ``.data``)."""
        self.__dict__.pop("_wver", None)

    def train(self, mode: bool = True):
        self.__dict__.pop("_wver", None)
        return super().train(mode)

    def _head_fwd(self, feats):
        K, dt = self._kernels_and_dtype()
        B, C, h, w = feats.shape
        ho, wo = self.adaptive_pool.output_size
        P, D = ho * wo, self.emb_dim
        ldo = (D + 7) // 8 * 8
        # packed copies of the 1x1 convolution (K-major weight, its transpose for the input gradient), refreshed when the
        # parameter changes
        ver = (self.conv1.weight._version, self.conv1.weight.data_ptr(), self.conv1.bias._version, dt)
        if self.__dict__.get("_wver") != ver:
            w2 = self.conv1.weight.detach().view(D, C)
            self.__dict__["_wpack"] = w2.to(dt).contiguous()
            wt = torch.zeros(C, ldo, dtype=dt, device=w2.device)
            wt[:, :D] = w2.t()
            self.__dict__["_wtpack"] = wt
            self.__dict__["_bpack"] = self.conv1.bias.detach().float().contiguous()
            self.__dict__["_wver"] = ver
        rows = torch.empty(B * P, C, dtype=dt, device=feats.device)
        K.pool_rows_fwd(feats.detach().float().contiguous(), rows, B, C, h, w, ho, wo)
        y = torch.zeros(B * P, ldo, dtype=dt, device=feats.device)
        K.gemm(rows, self.__dict__["_wpack"], y[:, :D], bias=self.__dict__["_bpack"])
        out = torch.empty(B, D, P, dtype=torch.float32, device=feats.device)
        K.pixels_bwd(y, out, B, D, P, P)
        return out, rows

    def _head_bwd(self, dout, rows, feat_shape, need_w, need_b, need_x):
        """dout (B, D, P) fp32 -> (d feats | None, d conv1.weight | None, d conv1.bias | None), all through the C ABI."""
        K, dt = self._kernels_and_dtype()
        B, C, h, w = feat_shape
        ho, wo = self.adaptive_pool.output_size
        P, D = ho * wo, self.emb_dim
        ldo = (D + 7) // 8 * 8
        dev = dout.device
        dy = torch.zeros(B * P, ldo, dtype=dt, device=dev)
        K.pixels_fwd(dout.contiguous().float(), dy, B, D, P, P)  # (B, D, P) -> rows (b, p) x D: the transpose of the hand-off
        dW = dB = dX = None
        if need_w or need_b:
            # flat fp32 gradient buffer [weight (D, C) | bias (D)], index maps as the decoder's packing maps
            key = (D, C, str(dev))
            if self.__dict__.get("_gmaps_key") != key:
                self.__dict__["_gmaps"] = (torch.arange(D, dtype=torch.int32, device=dev) * C, torch.arange(C, dtype=torch.int32, device=dev),
                                           torch.arange(D, dtype=torch.int32, device=dev) + D * C)
                self.__dict__["_gmaps_key"] = key
            rowoff, colmap, biasoff = self.__dict__["_gmaps"]
            g = torch.zeros(D * C + D, dtype=torch.float32, device=dev)
            K.wgrad(dy[:, :D], rows, g, rowoff, colmap, biasoff)
            if need_w:
                dW = g[: D * C].view(D, C, 1, 1)
            if need_b:
                dB = g[D * C :]
        if need_x:
            drows = torch.empty(B * P, C, dtype=dt, device=dev)
            K.gemm(dy, self.__dict__["_wtpack"], drows)  # d rows = d y @ W   (W^T stored K-major, pad columns zero)
            dX = torch.empty(B, C, h, w, dtype=torch.float32, device=dev)
            K.pool_rows_bwd(drows, dX, B, C, h, w, ho, wo)
        return dX, dW, dB

    def __getstate__(self):  # checkpoints pickle whole modules (G/utils.py:32-46): drop the device-side caches
        d = dict(self.__dict__)
        for k in ("_kernels", "_wpack", "_wtpack", "_bpack", "_wver", "_gmaps", "_gmaps_key"):
            d.pop(k, None)
        return d

    def fine_tune(self, fine_tune=True):
        for p in self.resnet.parameters():
            p.requires_grad = False
        for c in list(self.resnet.children())[5:]:
            for p in c.parameters():
                p.requires_grad = fine_tune


def _register(root: nn.Module, key: str, param: nn.Parameter) -> None:
    parts = key.split(".")
    m = root
    for name in parts[:-1]:
        if name not in m._modules:
            m.add_module(name, nn.Module())
        m = m._modules[name]
    m.register_parameter(parts[-1], param)


class _EncoderHeadFn(torch.autograd.Function):
    """Encoder hand-off under autograd: forward = the pooling / GEMM / transpose kernels, backward hand-written (Encoder._head_bwd)."""

    @staticmethod
    def forward(ctx, module, feats, weight, bias):
        out, rows = module._head_fwd(feats)
        ctx.module, ctx.rows, ctx.shape = module, rows, tuple(feats.shape)
        ctx.need = (feats.requires_grad, weight.requires_grad, bias.requires_grad)
        return out

    @staticmethod
    def backward(ctx, dout):
        need_x, need_w, need_b = ctx.need
        dX, dW, dB = ctx.module._head_bwd(dout, ctx.rows, ctx.shape, need_w, need_b, need_x)
        return None, dX, dW, dB


class _DecoderFn(torch.autograd.Function):
    """One autograd node for the whole decoder: forward = engine.forward, backward = the hand-written backward."""

    @staticmethod
    def forward(ctx, module, inp, seed, encoder_out, *params):
        eng = module._engine
        scores, ectx = eng.forward(inp, train=module.training, seed=seed)
        ctx.module, ctx.ectx = module, ectx
        ctx.need_enc = encoder_out.requires_grad
        ctx.param_req = [p.requires_grad for p in params]
        return scores

    @staticmethod
    def backward(ctx, dscores):
        module, ectx = ctx.module, ctx.ectx
        eng = module._engine
        B, T, W = dscores.shape
        ldW = (W + 7) // 8 * 8
        ds = torch.empty(B * T, ldW, dtype=eng.dtype, device=dscores.device)
        eng.K.cast2d(dscores.contiguous().view(B * T, W), ds, W)
        gflat = torch.zeros(eng.plan.n_params, dtype=torch.float32, device=dscores.device)
        d_enc = eng.backward(ectx, ds, gflat, need_encoder_grad=ctx.need_enc)
        if d_enc is not None:
            d_enc = d_enc[ectx.inp.unsort]  # back to the caller's (unsorted) batch order
        grads = []
        for (name, _), req in zip(eng.plan.shapes.items(), ctx.param_req):
            grads.append(eng.param(name, gflat) if req else None)
        return (None, None, None, d_enc, *grads)


class DecoderTransformer(nn.Module):
    """Generates the caption.  Subclasses fix ``variant`` ("G" geo-aware, "K" knowledge-aware, "N" news-knowledge-aware)."""

    variant = "G"
    # Test seam, set ONLY by tests/ (tests/hostsim.py checks the host-side orchestration on a GPU-less box).
    # The product never sets it: without it a non-CUDA device raises.
    _test_kernel_factory = None

    def __init__(self, word_map, emb_dim, decoder_dim, encoder_dim, num_heads, num_layers, dropout_dec=0.5, dropout_enc=0.5,
                 dropout_pos=0.1, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.word_map = word_map
        self.vocab_size = len(word_map)
        self.emb_dim = emb_dim
        self.decoder_dim, self.encoder_dim = decoder_dim, encoder_dim
        self.num_heads, self.num_layers = num_heads, num_layers
        self.dropout_dec, self.dropout_enc, self.dropout_pos = dropout_dec, dropout_enc, dropout_pos
        self.num_predicates = layout.NUM_PRED[self.variant]
        self.compute_dtype = compute_dtype
        self.lookahead_mask = None  # kept for attribute parity with the reference (the causal mask lives in the kernel)
        shapes = layout.param_shapes(self.variant, self.vocab_size, emb_dim, num_layers, decoder_dim, encoder_dim)
        n = sum(int(math.prod(s)) for s in shapes.values())
        flat = torch.zeros(n, dtype=torch.float32)
        off = 0
        self._param_names: List[str] = list(shapes.keys())
        for key, shp in shapes.items():
            cnt = int(math.prod(shp))
            p = nn.Parameter(flat[off : off + cnt].view(*shp))
            _register(self, key, p)
            off += cnt
        if self.variant != "G":
            # the reference registers one nn.Embedding under two names (K/models.py:330-331); keep both state_dict keys
            _register(self, "fact_encoder.predicate_embedding.weight", self._get("predicate_embedding.weight"))
        pe = torch.zeros(5000, emb_dim)
        position = torch.arange(0, 5000, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, emb_dim, 2).float() * (-math.log(10000.0) / emb_dim))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        if "pos_encoder" not in self._modules:
            self.add_module("pos_encoder", nn.Module())
        self.pos_encoder.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1))
        self._flat = flat
        self._flat_gen = 0  # bumped whenever the flat parameter buffer is re-gathered (captured graphs bake its addresses)
        self._engine: Optional[DecoderEngine] = None
        self._packed_version = None
        self._step = 0
        self.init_weights()

    # ---- parameter plumbing -------------------------------------------------------------------------------------------------
    def _get(self, key: str) -> nn.Parameter:
        m = self
        parts = key.split(".")
        for name in parts[:-1]:
            m = m._modules[name]
        return m._parameters[parts[-1]]

    def _params(self) -> List[nn.Parameter]:
        return [self._get(k) for k in self._param_names]

    def init_weights(self):
        """Same distributions as the reference's constructors (torch defaults) and init_weights (G/models.py:264-272)."""
        with torch.no_grad():
            for key in self._param_names:
                p = self._get(key)
                if key.endswith("in_proj_weight"):
                    nn.init.xavier_uniform_(p)
                elif key.endswith(("in_proj_bias", "out_proj.bias")) or (".norm" in key and key.endswith("bias")):
                    p.zero_()
                elif ".norm" in key and key.endswith("weight"):
                    p.fill_(1.0)
                elif key.endswith(("out_proj.weight", "linear1.weight", "linear2.weight")):
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                elif key.endswith(("linear1.bias", "linear2.bias")):
                    fan_in = self._get(key[:-4] + "weight").shape[1]
                    bound = 1 / math.sqrt(fan_in)
                    p.uniform_(-bound, bound)
                elif "embedding" in key:
                    p.normal_()
                elif key.startswith("fc_") and key.endswith("weight"):
                    p.uniform_(-0.1, 0.1)
                elif key.startswith("fc_") and key.endswith("bias"):
                    p.zero_()

    def load_pretrained_embeddings(self, embeddings):
        """G/models.py:274-280 (the reference swaps the Parameter; here the values are copied into the flat storage)."""
        with torch.no_grad():
            self._get("word_embedding.weight").copy_(embeddings)

    def fine_tune_embeddings(self, fine_tune=True):
        """G/models.py:282-289."""
        self._get("word_embedding.weight").requires_grad = fine_tune

    def repack(self) -> None:
        """
        Refresh the packed bf16 / transposed operand copies from the fp32 parameters NOW.  The module re-packs by itself when a
        parameter's autograd version changes (optimizer.step(), copy_, load_state_dict) and on every train() / eval()
        transition; writes that bypass the version counter - ``p.data.uniform_()`` (the reference's own idiom in init_weights,
        G/models.py:264-272), EMA or weight surgery through ``.data`` - need this call (or a train()/eval() switch) before the
        next forward.  Replaced storages (``p.data = ...``, ``load_state_dict(assign=True)``) are detected by address.
        """
        self._packed_version = None
        self._ensure_engine()

    invalidate = repack

    def train(self, mode: bool = True):
        # a cheap safety net for `.data` writes done between epochs: one re-pack launch per train()/eval() switch
        self._packed_version = None
        return super().train(mode)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"] = None  # ctypes handles and device buffers are rebuilt lazily after unpickling (G/utils.py:32-46)
        st.pop("_decode_graphs", None)
        st.pop("_decode_streams", None)
        st["_flat"] = None
        st["_packed_version"] = None
        return st

    # ---- engine -----------------------------------------------------------------------------------------------------------------
    def _ensure_engine(self) -> DecoderEngine:
        params = self._params()
        dev = params[0].device
        if dev.type != "cuda" and type(self)._test_kernel_factory is None:
            raise RuntimeError("ickb200 DecoderTransformer runs on a CUDA device only (no CPU fallback); call .to('cuda') first")
        flat = self._flat
        ok = flat is not None and flat.device == dev
        if ok:
            base, off = flat.data_ptr(), 0
            for p in params:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            # .to(device) / load_state_dict replaced the storages: gather into one flat buffer again and re-point the views
            n = sum(p.numel() for p in params)
            flat = torch.empty(n, dtype=torch.float32, device=dev)
            off = 0
            with torch.no_grad():
                for p in params:
                    cnt = p.numel()
                    flat[off : off + cnt].copy_(p.detach().reshape(-1).float())
                    p.data = flat[off : off + cnt].view(p.shape)
                    off += cnt
            self._flat = flat
            self._flat_gen = getattr(self, "_flat_gen", 0) + 1
            self._packed_version = None
            # captured decode loops (and the Trainer's captured steps, keyed on _flat_gen) read fp32 parameters - LayerNorm
            # weights, fc_entity, the type embedding - at addresses inside the OLD flat buffer
            self.__dict__.pop("_decode_graphs", None)
            if self._engine is not None and self._engine.device != dev:
                self._engine = None
        if self._engine is None:
            if type(self)._test_kernel_factory is not None:
                kernels = type(self)._test_kernel_factory()
            else:
                from .kernels import CudaKernels

                kernels = CudaKernels()
            plan = layout.PackPlan(self.variant, self.vocab_size, self.emb_dim, self.num_heads, self.num_layers, self.decoder_dim,
                                   self.encoder_dim)
            self._engine = DecoderEngine(plan, kernels, dev, self.compute_dtype or _default_dtype(), self.word_map["<pad>"],
                                         self.word_map["<start>"], self.word_map["<end>"], self.dropout_dec, self.dropout_enc,
                                         self.dropout_pos)
            self._packed_version = None
        self._engine.attach(self._flat)
        ver = (self._flat.data_ptr(), sum(p._version for p in params))
        if ver != self._packed_version:
            self._engine.repack()
            self._packed_version = ver
        return self._engine

    # ---- reference API --------------------------------------------------------------------------------------------------------------
    def _sorted_inputs(self, dev, captions, encoder_out, caption_masks, caption_lengths, entities, facts):
        # G/models.py:330.  stable=True: the reference's tie order is whatever torch.sort yields on its device; the CPU order
        # (which the golden vectors record) is the stable one, and CUDA's unstable sort would permute equal-length captions.
        lengths, sort_ind = caption_lengths.to(dev).squeeze(1).sort(dim=0, descending=True, stable=True)
        inp = NS()
        inp.captions = captions.to(dev)[sort_ind].contiguous()
        inp.caption_masks = caption_masks.to(dev)[sort_ind].contiguous()
        inp.encoder_out = encoder_out.detach().to(dev, torch.float32)[sort_ind].contiguous()
        inp.entities = entities.to(dev, torch.float32)[sort_ind].contiguous()  # arrives on the CPU (G/train.py:263-266)
        inp.facts = facts.to(dev)[sort_ind].contiguous() if facts is not None else None
        return inp, lengths, sort_ind

    def forward(self, captions, encoder_out, caption_masks, caption_lengths, entities, facts=None):
        if (self.variant == "G") != (facts is None):
            raise TypeError(f"variant {self.variant}: facts {'not accepted' if self.variant == 'G' else 'required'}")
        eng = self._ensure_engine()
        inp, lengths, sort_ind = self._sorted_inputs(eng.device, captions, encoder_out, caption_masks, caption_lengths, entities, facts)
        decode_lengths = (lengths - 1).tolist()
        seed = None
        if self.training:
            self._step += 1
            seed = (int(torch.initial_seed()) * 1000003 + self._step) & 0x7FFFFFFF
        params = self._params()
        if torch.is_grad_enabled() and (encoder_out.requires_grad or any(p.requires_grad for p in params)):
            # the autograd node needs the inverse permutation to hand d(encoder_out) back in the caller's order
            inp.unsort = torch.empty_like(sort_ind)
            inp.unsort[sort_ind] = torch.arange(sort_ind.numel(), device=sort_ind.device)
            scores = _DecoderFn.apply(self, inp, seed, encoder_out, *params)
        else:
            scores, _ = eng.forward(inp, train=self.training, seed=seed)
        return scores, inp.captions, decode_lengths

    def predict_batch(self, encoder_out, max_pred_len, entities, facts=None, return_margins=False):
        """Greedy decoding of a batch of images at once -> (B, max_pred_len) int64.  (The reference API is batch 1.)"""
        eng = self._ensure_engine()
        dev = eng.device
        inp = NS(encoder_out=encoder_out.detach().to(dev, torch.float32).contiguous(),
                 entities=entities.to(dev, torch.float32).contiguous(),
                 facts=facts.to(dev).contiguous() if facts is not None else None)
        graphed = (dev.type == "cuda" and not return_margins and type(self)._test_kernel_factory is None
                   and os.environ.get("ICKB200_DECODE_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing())
        with torch.no_grad():
            if not graphed:
                return eng.greedy_decode(inp, max_pred_len, return_margins=return_margins)
            return self._graphed_decode(eng, inp, max_pred_len)

    def beam_search_batch(self, encoder_out, max_pred_len, entities, facts=None, beam_size=5, return_scores=False):
        """
        EXTENSION (no reference counterpart: the reference's predict() is greedy): beam-search captions for a batch
        of images -> (B, max_pred_len) int64 in predict()'s token convention (no <start>, <end> included, <pad>-filled), and
        with return_scores the summed log-probability of each returned caption.  Device-resident, see engine.beam_decode.
        """
        eng = self._ensure_engine()
        dev = eng.device
        inp = NS(encoder_out=encoder_out.detach().to(dev, torch.float32).contiguous(),
                 entities=entities.to(dev, torch.float32).contiguous(),
                 facts=facts.to(dev).contiguous() if facts is not None else None)
        graphed = (dev.type == "cuda" and type(self)._test_kernel_factory is None
                   and os.environ.get("ICKB200_DECODE_GRAPH", "1") != "0" and not torch.cuda.is_current_stream_capturing())
        with torch.no_grad():
            if graphed:
                out, score = self._graphed_decode(eng, inp, max_pred_len, beam=int(beam_size))
            else:
                out, score = eng.beam_decode(inp, max_pred_len, int(beam_size))
        return (out, score) if return_scores else out

    def _graphed_decode(self, eng, inp, max_pred_len, beam=0):
        """
        The whole decode loop (context encoders, memory K/V, max_pred_len x ~40 kernels, all device-resident) as ONE CUDA graph
        per input shape: launched eagerly from Python the loop is bound by the host (~20 us per launch), replayed it is bound by
        the GPU.  Weights are read from the packed operand buffers at fixed addresses, so parameter updates are picked up by the
        usual re-pack; inputs are copied into the graph's static buffers, the tokens are copied out.
        """
        cache = self.__dict__.setdefault("_decode_graphs", {})
        key = (id(eng), max_pred_len, tuple(inp.encoder_out.shape), tuple(inp.entities.shape),
               tuple(inp.facts.shape) if inp.facts is not None else None, beam)
        one = (lambda x: eng.beam_decode(x, max_pred_len, beam)) if beam else (lambda x: eng.greedy_decode(x, max_pred_len))
        # Experiment switch, default 1 = off: images are independent, so the batch can be cut into ICKB200_DECODE_STREAMS chunks whose
        # loops are captured on parallel branches of the graph (one chunk's small latency-bound GEMMs / LayerNorms could overlap
        # another chunk's HBM-bound attention).  Measured on a B200: 23.3k captions/s with 1, 22.2k / 21.3k / 19.9k with 2 / 3 / 4
        # branches - the halves' kernels do not overlap enough to pay for running twice as many of them (profiles/README.md).
        nstr = max(1, min(int(os.environ.get("ICKB200_DECODE_STREAMS", "1")), inp.encoder_out.shape[0] // 64))
        side_streams = self.__dict__.setdefault("_decode_streams", [])
        while len(side_streams) < nstr:
            side_streams.append(torch.cuda.Stream())

        def run(x):
            if nstr == 1:
                return one(x)
            B = x.encoder_out.shape[0]
            cuts = [B * i // nstr for i in range(nstr + 1)]
            cur = torch.cuda.current_stream()
            outs = []
            for i in range(nstr):
                st = side_streams[i]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    sl = slice(cuts[i], cuts[i + 1])
                    outs.append(one(NS(encoder_out=x.encoder_out[sl], entities=x.entities[sl],
                                       facts=x.facts[sl] if x.facts is not None else None)))
            for i in range(nstr):
                cur.wait_stream(side_streams[i])
            if beam:
                return tuple(torch.cat([o[j] for o in outs]) for j in range(2))
            return torch.cat(outs)

        key = key + (nstr,)
        entry = cache.get(key)
        if entry is None:
            static = NS(encoder_out=inp.encoder_out.clone(), entities=inp.entities.clone(),
                        facts=inp.facts.clone() if inp.facts is not None else None)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # eager warm-up: one-time initialisation must not happen during capture
                run(static)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run(static)
            if len(cache) >= 4:  # each graph owns its activation pool: keep a few shapes only
                cache.pop(next(iter(cache)))
            entry = cache[key] = (graph, static, out)
        graph, static, out = entry
        static.encoder_out.copy_(inp.encoder_out, non_blocking=True)
        static.entities.copy_(inp.entities, non_blocking=True)
        if static.facts is not None:
            static.facts.copy_(inp.facts, non_blocking=True)
        graph.replay()
        return tuple(o.clone() for o in out) if beam else out.clone()

    def predict(self, encoder_out, max_pred_len, entities, facts=None):
        """G/models.py:363-443: batch-1 greedy decode -> (max_pred_len, 1) int64."""
        out = self.predict_batch(encoder_out, max_pred_len, entities, facts)
        return out.transpose(0, 1).contiguous()
