"""
B200-native caption-decoder hot path of sonniki/image-captioning-with-external-knowledge.

Drop-in behind the reference's own module API (``models.Encoder`` / ``models.DecoderTransformer.forward`` /
``.predict`` for the geo-aware, knowledge-aware and news-knowledge-aware variants).  All arithmetic runs in
hand-written sm_100a CUDA kernels reached through the C-ABI library ``csrc/libickb200.so`` (declared in
``include/ickb200.h``); there is no CPU fallback — importing the kernels without the built library raises.
"""
__all__ = ["synthetic", "layout", "kernels", "engine", "models", "trainer"]
