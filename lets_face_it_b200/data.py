"""Input side of the path on the device.

The reference's `MimicryDataset` (mimicry_data_module.py:12-81) enumerates every stride-1 window of `seq_len` frames of
every segment and, per item, re-reads the window's frames of up to four modalities from HDF5 — consecutive items of a
segment share `seq_len - 1` of their `seq_len` frames, so a batch of 256 windows moves 14 MB host -> device to deliver
0.2 MB of new frames.  `ResidentWindows` keeps the corpus (11.5 h of 25 fps features = 2.07 M frames x 172 floats =
1.4 GB: < 1 % of a B200's HBM) on the device once, reproduces the reference's window table (same enumeration order, same
`random.sample` shuffle), and builds a batch with ONE small host -> device copy (the window start rows, 8 bytes per
sequence) and one gather launch per modality (`lfi_gather_batch`).  The batch dict it returns is what
`SeqGlow.forward` / `Trainer.step` take (`p1_face, p2_face [B,T,56]`, `p1_speech, p2_speech [B,T,30]`).

HDF5 itself is outside this package (h5py is not available offline): `from_segments` takes the per-segment arrays the
reference reads (`/{split}/{flame_expression,flame_jaw,flame_neck,mfcc,prosody}/{segment}/{agent,interlocutor}`,
combine_features.py:214-216) already assembled per modality as `MimicryDataset.__getitem__` assembles them
(face = expression[:, :expression_dim] | jaw | neck, speech = mfcc | prosody; :55-66).
"""
from __future__ import annotations

import random
from typing import Dict, List, Sequence

import torch

from . import _cabi as cabi

MODALITIES = ("p1_face", "p2_face", "p1_speech", "p2_speech")


def window_table(lengths: Sequence[int], seq_len: int, shuffle=True, rng=None) -> List[int]:
    """Global first rows of the reference's window table (mimicry_data_module.py:35-42): for every segment with at least
    `seq_len` frames every stride-1 window, in segment order, then `random.sample` over the whole table (the reference's
    shuffle: same RNG consumption).  Row = segment offset in the concatenated corpus + window start."""
    table: List[int] = []
    off = 0
    for L in lengths:
        L = int(L)
        if L >= seq_len:
            table.extend(off + s for s in range(L - seq_len + 1))
        off += L
    if shuffle:
        table = (rng or random).sample(table, len(table))
    return table


class ResidentWindows:
    def __init__(self, segments: Sequence[Dict[str, torch.Tensor]], seq_len: int, device, shuffle=True, rng=None):
        """segments: one dict per segment, modality -> [L_i, dim] float tensor (all modalities of a segment share L_i)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("lets_face_it_b200.data: the resident corpus lives on a CUDA device (got %s); there is no CPU path" % dev)
        self.device, self.seq_len = dev, int(seq_len)
        mods = [m for m in MODALITIES if m in segments[0]]
        self.raw, self.dim = {}, {}
        offs, o = [], 0
        for seg in segments:
            offs.append(o)
            o += int(seg[mods[0]].shape[0])
        self.rows = o
        for m in mods:
            self.raw[m] = torch.cat([seg[m].to(torch.float32) for seg in segments], dim=0).contiguous().to(dev)
            self.dim[m] = int(self.raw[m].shape[1])
        self.table = torch.tensor(window_table([int(seg[mods[0]].shape[0]) for seg in segments], self.seq_len, shuffle, rng), dtype=torch.int64)
        self._pin = None
        self._dev_idx = None

    def __len__(self):
        return int(self.table.numel())

    def batch(self, index: Sequence[int] | torch.Tensor, out: Dict[str, torch.Tensor] | None = None):
        """The batch the DataLoader would collate from items `index` of the table: {modality: [B, seq_len, dim]} on the device."""
        idx = torch.as_tensor(index, dtype=torch.int64)
        B = int(idx.numel())
        if self._pin is None or self._pin.numel() < B:
            self._pin = torch.empty(B, dtype=torch.int64).pin_memory()
            self._dev_idx = torch.empty(B, dtype=torch.int64, device=self.device)
        torch.index_select(self.table, 0, idx, out=self._pin[:B])
        self._dev_idx[:B].copy_(self._pin[:B], non_blocking=True)   # the step's whole host -> device traffic: 8 bytes per sequence
        L = cabi.lib()
        res = out if out is not None else {}
        with torch.cuda.device(self.device):
            st = cabi.stream_ptr()
            for m, raw in self.raw.items():
                t = res.get(m)
                if t is None or tuple(t.shape) != (B, self.seq_len, self.dim[m]):
                    t = torch.empty(B, self.seq_len, self.dim[m], dtype=torch.float32, device=self.device)
                    res[m] = t
                cabi.check(L.lfi_gather_batch(raw.data_ptr(), self._dev_idx.data_ptr(), B, self.seq_len, self.dim[m], t.data_ptr(), st),
                           "lfi_gather_batch")
        return res

    @property
    def h2d_bytes_per_sequence(self):
        return 8
