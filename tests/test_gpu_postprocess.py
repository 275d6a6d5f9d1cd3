"""GPU tests of the SURVEY.md §8(f) rows: the output wire format of sampling (`lfi_expand_faces`: de-standardise +
56 -> 106 FLAME scatter in one launch) bit-exact against the reference-generated vectors and the oracle, and the
training-step glue (`Trainer.training_step`: mismatched-NLL probe with the -0.1 loss scale) of the fused step."""
import random

import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from tests.helpers import final_hparams, load_golden
from tests.kat import build_kat_model, kat_batch, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_expand_faces_bit_exact():
    from lets_face_it_b200.postprocess import destandardize_expand, expand_face_dim

    g = load_golden("kat_post")
    hp = final_hparams()
    seq, means, stds = (torch.from_numpy(np.asarray(g[k])).to(DEV) for k in ("seq", "means", "stds"))
    assert torch.equal(expand_face_dim(seq, hp.Data).cpu(), torch.from_numpy(np.asarray(g["expanded"])))
    assert torch.equal(destandardize_expand(seq, means, stds, hp.Data).cpu(), torch.from_numpy(np.asarray(g["destd_expanded"])))
    # BASELINE configs[3] shape per GPU: 1024 sequences x 750 frames, against the oracle
    e, j, n = hp.Data["expression_dim"], hp.Data["jaw_dim"], hp.Data["neck_dim"]
    big = torch.randn(1024, 750, 56, generator=torch.Generator().manual_seed(5))
    out = destandardize_expand(big.to(DEV), means, stds, hp.Data).cpu()
    assert torch.equal(out, O.destandardize_expand(big, means.cpu(), stds.cpu(), e, j, n))
    with pytest.raises(RuntimeError):
        expand_face_dim(big[:2], hp.Data)  # CPU tensor: no fallback


def test_training_step_probe_and_loss_scale():
    """lets_face_it_glow.py:39-55: the probe fires only when use_negative_nll_loss, last mismatched NLL > 0 and
    random() < 0.1 (evaluated in that order); its gradient is -0.1 x the gradient of the deranged batch."""
    from lets_face_it_b200.train import Trainer

    hp = final_hparams()
    m = build_kat_model(hp, DEV)
    m.glow.set_actnorm_init(True)
    m.train()
    batch = to_device(kat_batch(hp, 16, 40, seed=31), DEV)
    tr = Trainer(m, dropout=False, lr=0.0)
    assert tr.missmatched_modalities == ["p2_face", "p2_speech"] and tr.last_missmatched_nll == float("inf")
    tr.step(batch, loss_scale=1.0)
    g1 = tr.gflat.clone()
    tr.step(batch, loss_scale=-0.1)
    g2 = tr.gflat.clone()
    n = tr.eng.n_theta
    assert float((g2[:n] + 0.1 * g1[:n]).norm() / (0.1 * g1[:n].norm())) < 1e-5
    # probe schedule: python's random stream decides, consumed only while the last mismatched NLL is positive
    random.seed(4)
    draws = [random.random() for _ in range(40)]
    random.seed(4)
    fired = []
    for i in range(40):
        loss, deranged = tr.training_step(batch)
        fired.append(deranged)
        if deranged:
            assert tr.last_missmatched_nll == pytest.approx(-float(loss) / -0.1, rel=1e-6)
        if tr.last_missmatched_nll <= 0:
            break
    k = len(fired)
    assert fired == [d < 0.1 for d in draws[:k]]
    tr.last_missmatched_nll = -1.0  # a non-positive probe NLL switches the probe off without consuming random numbers
    state = random.getstate()
    _, deranged = tr.training_step(batch)
    assert not deranged and random.getstate() == state


def test_host_feed_matches_device_steps():
    """`HostFeed` (pinned host batch -> copy stream -> step, loss read back one step late) returns exactly the losses of
    `Trainer.step` on device-resident copies of the same batches, in order."""
    from lets_face_it_b200.train import HostFeed, Trainer

    hp = final_hparams()
    m = build_kat_model(hp, DEV)
    m.glow.set_actnorm_init(True)
    m.train()
    tr = Trainer(m, dropout=False, lr=0.0)
    hosts = [{k: v.pin_memory() for k, v in kat_batch(hp, 16, 40, seed=40 + i).items()} for i in range(4)]
    want = [float(tr.step(to_device(h, DEV))) for h in hosts]
    feed = HostFeed(tr)
    got = [feed.step(h) for h in hosts] + [feed.flush()]
    assert got[0] is None and feed.flush() is None
    assert got[1:] == pytest.approx(want, rel=1e-6)


def test_calc_jerk_on_device_matches_reference_vector():
    """`calc_jerk` (glow/utils.py:53-58; the validation metric of mimicry_logger.py:175-184) as one device launch against the
    value the reference computed on the same frames (kat_post.npz), and at the sampling size 1024 x 750 against torch."""
    from lets_face_it_b200.glow.utils import calc_jerk

    g = load_golden("kat_post")
    x = torch.from_numpy(np.asarray(g["jerk_x"])).to(DEV)
    got = calc_jerk(x)
    assert got.is_cuda and abs(float(got) - float(g["jerk"])) <= 2e-7 * float(g["jerk"])
    big = torch.randn(1024, 750, 56, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    d = big[:, 1:] - big[:, :-1]
    a = d[:, 1:] - d[:, :-1]
    ref = (a[:, 1:] - a[:, :-1]).abs().double().mean()
    assert abs(float(calc_jerk(big)) - float(ref)) <= 1e-6 * float(ref)
