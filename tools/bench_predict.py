"""
Greedy caption generation throughput (BASELINE.json configs[3]: 5k images over 8 GPUs = 625 images per GPU;
the reference has no beam search, SURVEY.md §0 — the pinned mode is greedy predict() + repetition clean-up).
    python tools/bench_predict.py [--variant K] [--images 625] [--dtype bf16]
Prints one JSON line: images/s on one GPU through DecoderTransformer.predict_batch (host inputs, D2H of the tokens),
plus the CPU port of predict() timed on one image (batch 1, as the reference requires).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ickb200  # noqa
from ickb200 import layout, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="K")
ap.add_argument("--images", type=int, default=625)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--beam", type=int, default=0, help="0 = greedy predict(); k > 0 = beam search of width k (extension, no reference counterpart)")
a = ap.parse_args()
name = {"G": "geo_b32", "K": "knowledge_b128", "N": "news_b8"}[a.variant]
cfg = syn.BASELINE_CONFIGS[name].with_batch(a.images)
Tmax = 30 if a.variant == "G" else 40
mod = {"G": "geo_aware", "K": "knowledge_aware", "N": "news_knowledge_aware"}[a.variant]
DecoderTransformer = __import__(f"ickb200.{mod}", fromlist=["DecoderTransformer"]).DecoderTransformer
torch.manual_seed(0)
dec = DecoderTransformer(syn.make_word_map(cfg.V), cfg.D, cfg.ff, cfg.ff, cfg.H, cfg.L,
                         compute_dtype={"bf16": torch.bfloat16, "fp32": torch.float32}[a.dtype])
shapes = layout.param_shapes(cfg.variant, cfg.V, cfg.D, cfg.L, cfg.ff, cfg.ff)
w = syn.det_weights(shapes)
dec.load_state_dict(w, strict=False)
dec = dec.cuda().eval()
b = syn.make_batch(cfg, seed=7)
enc, ent, facts = b["encoder_out"].pin_memory(), b["entities"], b.get("facts")


def run():
    if a.beam:
        out = dec.beam_search_batch(enc.cuda(non_blocking=True), Tmax, ent, facts.cuda() if facts is not None else None, beam_size=a.beam)
    else:
        out = dec.predict_batch(enc.cuda(non_blocking=True), Tmax, ent, facts.cuda() if facts is not None else None)
    return out.cpu()


run()
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    out = run()
    ts.append(time.perf_counter() - t0)
sec = sorted(ts)[len(ts) // 2]
lens = (out != 0).sum(1).float()
# CPU port, one image (the reference's predict is batch 1)
from oracle import decoder_oracle as orc

p = dict(w)
p["pos_encoder.pe"] = orc.positional_table(5000, cfg.D).unsqueeze(1)
spec = orc.Spec(cfg.variant, cfg.V, cfg.D, cfg.H, cfg.L, pad=0, start=cfg.V - 2, end=cfg.V - 1)
torch.set_num_threads(os.cpu_count())
t0 = time.perf_counter()
with torch.no_grad():
    if a.beam:
        ref = orc.beam_search(spec, p, b["encoder_out"][:1], Tmax, ent[:1], facts[:1] if facts is not None else None, beam_size=a.beam)
    else:
        ref = orc.predict(spec, p, b["encoder_out"][:1], Tmax, ent[:1], facts[:1] if facts is not None else None)
cpu_sec = time.perf_counter() - t0
print(json.dumps({"metric": f"beam{a.beam}_captions_per_sec" if a.beam else "greedy_captions_per_sec", "value": a.images / sec, "unit": "captions/s", "n_gpus": 1, "variant": a.variant,
                  "images": a.images, "max_len": Tmax, "dtype": a.dtype, "sec": sec, "mean_generated_len": float(lens.mean()),
                  "first_image_matches_cpu_port": out[0].tolist() == ref.reshape(-1).tolist(),
                  "cpu_port": {"captions_per_sec": 1.0 / cpu_sec, "cores": os.cpu_count(), "sample": "1 image, batch 1, full re-decode per step"}}))
