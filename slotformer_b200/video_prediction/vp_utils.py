"""Evaluation helper of the video-prediction task that sits right after the decoder.

``postproc_mask`` mirrors reference slotformer/video_prediction/vp_utils.py:20-41 (same name, argument and
result) and runs on the sm_100a kernel (csrc/decode_combine.cu); like the other operators of this package it
has no CPU path (engine.SfbError for non-CUDA tensors).
"""
from ..engine import FG_THRE, postproc_mask as _postproc_mask_kernel


def postproc_mask(batch_masks):
    """batch_masks [B, T, N, 1, H, W] float32 (CUDA) -> masks [B, T, H, W] int64, slot index per pixel.

    The slot whose largest mask value is smallest is the frame's background; pixels whose best score is below
    FG_THRE are given to it, every other pixel to its arg-max slot (first maximum)."""
    return _postproc_mask_kernel(batch_masks, FG_THRE)
