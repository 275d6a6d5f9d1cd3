"""Output side of sampling: the step after `SeqGlow.inference`.

Mirrors `generate_motion_from_model.py` of the reference: `expand_face_dim(seq, data_hparams)` (:39-51) scatters the
56 generated channels (50 expression + 3 jaw + 3 neck) into the 106-wide FLAME vector the render server consumes, and
`generate_motion` (:54-70) de-standardises first (`predicted_seq * face_stds + face_means`, :68).  Both run as ONE
coalesced CUDA launch through the C ABI (`lfi_expand_faces`); CPU tensors raise (no fallback).
"""
from __future__ import annotations

import torch

from . import _cabi as cabi


def _dims(data_hparams):
    return int(data_hparams["expression_dim"]), int(data_hparams["jaw_dim"]), int(data_hparams["neck_dim"])


def destandardize_expand(seq, face_means, face_stds, data_hparams):
    """[B, T, C] standardised frames -> [B, T, 106] FLAME vectors (generate_motion_from_model.py:68 then :39-51)."""
    if seq.device.type != "cuda":
        raise RuntimeError("lets_face_it_b200.postprocess: frames live on %s; the path runs on CUDA only (no CPU path)" % seq.device)
    e, j, n = _dims(data_hparams)
    if seq.dim() != 3 or seq.shape[2] < e + j + n:
        raise RuntimeError("expected [B, T, >=%d] frames, got %s" % (e + j + n, tuple(seq.shape)))
    x = seq[:, :, :e + j + n].to(torch.float32).contiguous()
    out = torch.empty(x.shape[0], x.shape[1], 106, dtype=torch.float32, device=x.device)
    m = s = None
    if face_means is not None:
        m = face_means.to(device=x.device, dtype=torch.float32).reshape(-1)[:e + j + n].contiguous()
        s = face_stds.to(device=x.device, dtype=torch.float32).reshape(-1)[:e + j + n].contiguous()
    cabi.check(cabi.lib().lfi_expand_faces(x.data_ptr(), cabi.ptr(m), cabi.ptr(s), x.shape[0] * x.shape[1], e, j, n, out.data_ptr(),
                                           cabi.stream_ptr()), "lfi_expand_faces")
    return out


def expand_face_dim(seq, data_hparams):
    """Same signature as the reference's `expand_face_dim` (generate_motion_from_model.py:39-51; mimicry_logger.py:49-63)."""
    return destandardize_expand(seq, None, None, data_hparams)
