"""Drop-in for the reference's ``news-knowledge-aware/models.py``: ``from ickb200.news_knowledge_aware import *``."""
from . import models as _m
from .models import Encoder, device  # noqa: F401


class DecoderTransformer(_m.DecoderTransformer):
    variant = "N"


__all__ = ["Encoder", "DecoderTransformer", "device"]
