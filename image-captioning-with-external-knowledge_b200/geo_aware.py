"""Drop-in for the reference's ``geo-aware/models.py``: ``from ickb200.geo_aware import *``."""
from . import models as _m
from .models import Encoder, device  # noqa: F401


class DecoderTransformer(_m.DecoderTransformer):
    variant = "G"


__all__ = ["Encoder", "DecoderTransformer", "device"]
