"""Raw-feature ingest on the device — the step right before the hot path.

Reference (host, numpy/torch, one video at a time): ``video_nmn/dataset.py:134-172``
  * RX / TGIF-QA h5 features: ``resnet_features[id]`` [clips, frames, 2048] -> ``[:max_video_length]`` -> ``mean(dim=1)``;
    ``resnext_features[id]`` [clips, 2048] -> ``[:max_video_length]`` -> ``torch.cat([appearance, motion], dim=-1)``   (:145-172)
  * I3D npy features: rows ``np.arange(0, n, 2)`` -> ``[:max_video_length]``                                          (:134-143)

Here the same reductions run batched in HBM-bound CUDA kernels (csrc/ingest.cu) and write the model's input ``[B, T, V]``
directly in its storage dtype.  No CPU path: tensors must be on a CUDA device.
"""
from __future__ import annotations

import torch

from . import _lib as L


def pool_concat(appearance: torch.Tensor, motion: torch.Tensor | None = None, out_dtype=torch.bfloat16, max_video_length=None,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """appearance [B, T, F, Da] (+ motion [B, T, Dm]) -> video features [B, T', Da + Dm], T' = min(T, max_video_length)."""
    L.require_cuda(appearance, 'appearance')
    if appearance.dim() != 4:
        raise ValueError('appearance must be [B, clips, frames, D], got %s' % (tuple(appearance.shape),))
    B, T, F, Da = appearance.shape
    Dm = 0
    if motion is not None:
        L.require_cuda(motion, 'motion')
        if motion.dim() != 3 or motion.shape[0] != B or motion.shape[1] != T or motion.dtype != appearance.dtype:
            raise ValueError('motion must be [B, clips, D] with the dtype of appearance')
        Dm = motion.shape[2]
    if max_video_length is not None and T > max_video_length:          # dataset.py:148-149,166-167 truncate the clips first
        appearance = appearance[:, :max_video_length]
        motion = motion[:, :max_video_length] if motion is not None else None
        T = max_video_length
        if B > 1:                                                        # the kernel wants dense [B, T, ...]
            appearance = appearance.contiguous()
            motion = motion.contiguous() if motion is not None else None
    appearance = appearance.contiguous()
    motion = motion.contiguous() if motion is not None else None
    if out is None:
        out = torch.empty((B, T, Da + Dm), dtype=out_dtype, device=appearance.device)
    rc = L.lib().stair_ingest_pool_concat(L.ptr(appearance), L.ptr(motion), L.i32(L.dtype_code(appearance.dtype)), L.ptr(out),
                                          L.i32(L.dtype_code(out.dtype)), L.i32(B), L.i32(T), L.i32(F), L.i32(Da), L.i32(Dm), L.stream_ptr())
    L.check(rc, 'stair_ingest_pool_concat')
    return out


def subsample(feats: torch.Tensor, max_video_length: int, step: int = 2, out_dtype=torch.bfloat16) -> torch.Tensor:
    """feats [B, n, D] -> [B, min(ceil(n / step), max_video_length), D] with rows 0, step, 2 step, ... (dataset.py:138-141)."""
    L.require_cuda(feats, 'feats')
    if feats.dim() != 3:
        raise ValueError('feats must be [B, frames, D]')
    B, n, D = feats.shape
    T = min((n + step - 1) // step, max_video_length)
    feats = feats.contiguous()
    out = torch.empty((B, T, D), dtype=out_dtype, device=feats.device)
    rc = L.lib().stair_ingest_subsample(L.ptr(feats), L.i32(L.dtype_code(feats.dtype)), L.ptr(out), L.i32(L.dtype_code(out.dtype)), L.i32(B),
                                        L.i32(n), L.i32(T), L.i32(D), L.i32(step), L.stream_ptr())
    L.check(rc, 'stair_ingest_subsample')
    return out
