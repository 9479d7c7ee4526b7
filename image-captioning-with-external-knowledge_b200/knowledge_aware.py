"""Drop-in for the reference's ``knowledge-aware/models.py``: ``from ickb200.knowledge_aware import *``."""
from . import models as _m
from .models import Encoder, device  # noqa: F401


class DecoderTransformer(_m.DecoderTransformer):
    variant = "K"


__all__ = ["Encoder", "DecoderTransformer", "device"]
