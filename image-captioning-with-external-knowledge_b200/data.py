"""
Input pipeline, device half.

The reference's ``CaptionDataset`` reads images from HDF5 as fp16 ``(3, 256, 256)`` arrays with values in [0, 255]
(G/create_input_files.py:99-101, :334-337), divides by 255 on the host (``imgs[i] / 255.`` — a numpy fp16 division — then
``torch.FloatTensor``, G/datasets.py:44), applies ``transforms.Normalize(IMAGENET_MEAN, IMAGENET_STD)`` per image in a
DataLoader worker (G/train.py:139-147) and copies the fp32 batch to the device (G/train.py:263): 4 bytes per element over
PCIe and a Python call per image.  Here the batch crosses PCIe once in its 2-byte storage format and one HBM-bound kernel
(``ick_image_prep``) reproduces the same arithmetic — same roundings, the fp32 result is bit-identical — writing the layout and
dtype the encoder trunk wants.  The HDF5 / JSON / pickle readers themselves are host file I/O and are not part of this package.
"""
from __future__ import annotations

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # G/train.py:139-141
IMAGENET_STD = (0.229, 0.224, 0.225)


def prepare_images(raw: torch.Tensor, dtype: torch.dtype = torch.float32, channels_last: bool = False, mean=IMAGENET_MEAN,
                   std=IMAGENET_STD, kernels=None) -> torch.Tensor:
    """
    raw: (N, 3, H, W) fp16 CUDA tensor holding the stored pixel values (0..255).  Returns the tensor ``train.py`` would feed to
    ``Encoder.forward`` — ``Normalize(mean, std)(FloatTensor(raw / 255.))`` — as ``dtype`` (fp32 or bf16), logically (N, 3, H, W),
    in contiguous or channels-last memory format.  CUDA only: there is no CPU fallback.
    """
    if not raw.is_cuda:
        raise RuntimeError("ickb200.data.prepare_images: CUDA tensor required (the kernels have no CPU fallback)")
    if raw.dtype != torch.float16:
        raise TypeError("prepare_images expects the fp16 storage format of the reference's HDF5 files")
    if kernels is None:
        from .kernels import CudaKernels

        kernels = CudaKernels()
    raw = raw.contiguous()
    N, C, H, W = raw.shape
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    out = torch.empty((N, C, H, W), dtype=dtype, device=raw.device, memory_format=fmt)
    kernels.image_prep(raw, out, mean, std, channels_last=channels_last)
    return out
