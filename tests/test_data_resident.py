"""Input side on the device: the HBM-resident window batching against the oracle's restatement of
MimicryDataset (window table incl. the random.sample shuffle, __getitem__, default collate)."""
import random

import pytest
import torch

from oracle import glow_oracle as O


def _segments(lengths, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"p1_face": torch.randn(L, 56, generator=g), "p2_face": torch.randn(L, 56, generator=g),
             "p1_speech": torch.randn(L, 30, generator=g), "p2_speech": torch.randn(L, 30, generator=g)} for L in lengths]


def test_window_table_matches_reference_enumeration_and_shuffle():
    from lets_face_it_b200.data import window_table

    lengths = [100, 79, 80, 231, 5, 81]   # 79 and 5 are shorter than seq_len: skipped, as the reference skips them
    segs = _segments(lengths)
    ref = O.dataset_window_table(segs, 80, random.Random(5))
    got = window_table(lengths, 80, True, random.Random(5))
    offs = [0]
    for L in lengths:
        offs.append(offs[-1] + L)
    assert len(got) == len(ref) == (21 + 1 + 152 + 2)
    assert got == [offs[k] + idx[0] for k, idx in ref]
    assert all(idx == list(range(idx[0], idx[0] + 80)) for _, idx in ref)
    assert window_table(lengths, 80, False) == sorted(got)
    assert window_table([10, 20], 80, True) == []   # nothing long enough: empty table


def test_resident_corpus_has_no_cpu_path():
    from lets_face_it_b200.data import ResidentWindows

    with pytest.raises(RuntimeError):
        ResidentWindows(_segments([90]), 80, "cpu")


@pytest.mark.gpu
def test_resident_batches_equal_dataloader_batches():
    """Batches of 256 windows (and a ragged last batch of 17) out of the resident corpus are bit-identical to what the
    reference's Dataset + DataLoader collate would deliver for the same shuffled table, and feed SeqGlow.forward directly."""
    from lets_face_it_b200.data import ResidentWindows

    lengths = [300, 95, 80, 1200, 40, 777]
    segs = _segments(lengths, seed=3)
    ds = ResidentWindows(segs, 80, "cuda:0", shuffle=True, rng=random.Random(11))
    table = O.dataset_window_table(segs, 80, random.Random(11))
    assert len(ds) == len(table)
    for index in (list(range(256)), list(range(len(ds) - 17, len(ds))), [5, 5, 0, len(ds) - 1]):
        got = ds.batch(index)
        ref = O.dataset_batch(segs, table, index)
        for m in ref:
            assert got[m].shape == ref[m].shape
            assert torch.equal(got[m].cpu(), ref[m]), m
    assert ds.h2d_bytes_per_sequence == 8
