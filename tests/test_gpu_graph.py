"""The whole training step as ONE CUDA graph (lets_face_it_glow.py:39-72): `GraphedStep` replays the
launch sequence of `Trainer.step` - and must train exactly as the eager step does."""
import pytest
import torch

from tests.helpers import final_hparams, small_hparams
from tests.kat import build_kat_model, kat_batch, to_device

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def _mode(name):
    from lets_face_it_b200 import _cabi as cabi

    return {"fp32": cabi.GEMM_FP32, "bf16x3": cabi.GEMM_BF16X3, "bf16": cabi.GEMM_BF16}[name]


@pytest.mark.parametrize("case,mode", [("small", "fp32"), ("full", "bf16x3")])
def test_graphed_steps_equal_eager_steps(case, mode):
    """Same model, same batches: 2 eager steps + 4 graph replays against 6 eager steps.  Losses of every step and the final
    parameters agree to the run-to-run noise of the atomically accumulated gradients (the launches are the same ones)."""
    from lets_face_it_b200.train import Trainer

    hp = small_hparams("gru") if case == "small" else final_hparams()
    B, T = (6, 30) if case == "small" else (32, 40)
    batches = [to_device(kat_batch(hp, B, T, seed=60 + i), DEV) for i in range(6)]
    runs = []
    for graphed in (False, True):
        m = build_kat_model(hp, DEV).train()
        m.gemm_mode = _mode(mode)
        tr = Trainer(m, dropout=False)
        tr.lr = 1e-3  # large enough that six steps move the loss visibly
        losses = []
        if not graphed:
            for b in batches:
                losses.append(float(tr.step(b)))
        else:
            for b in batches[:1]:
                losses.append(float(tr.step(b)))
            gs = tr.graphed(batches[1], warmup=1)       # one more eager step on batches[1] inside, then the capture
            losses.append(None)
            for b in batches[2:]:
                losses.append(float(gs.step(b)))
            assert tr.step_count == 6
        runs.append((losses, {n: p.detach().clone() for n, p in m.named_parameters()}))
    (l0, p0), (l1, p1) = runs
    for st in (0, 2, 3, 4, 5):
        assert abs(l0[st] - l1[st]) < 2e-5 * abs(l0[st]) + 1e-6, (st, l0[st], l1[st])
    assert abs(l0[5] - l0[0]) > 1e-3 * abs(l0[0])  # the steps did train
    # Adam's first updates are ~ lr * sign(g): elements whose gradient is at rounding level flip with the order of the atomic
    # accumulations from run to run, eager or graphed - hence a bound on the update (6 steps x lr 1e-3), not bit equality
    for n in p0:
        d = (p0[n] - p1[n]).double().norm() / p0[n].double().norm().clamp_min(1e-12)
        assert float(d) < 1e-3, (n, float(d))


def test_graphed_step_with_frame_dropout_draws_fresh_masks():
    """With the frame dropout of final_model.yaml active, every replay must draw NEW masks (torch advances the captured
    generator per replay): the same batch replayed twice gives different losses, as two eager steps do."""
    from lets_face_it_b200.train import Trainer

    hp = final_hparams()
    m = build_kat_model(hp, DEV).train()
    for name in ("p2_face", "p1_speech", "p2_speech"):
        enc = getattr(m.feature_encoder, name + "_encoder", None)
        pdrop = hp.Conditioning[name]["dropout"]
        if enc is not None and pdrop > 0:  # build_kat_model disables the frame dropout for parity runs
            enc.dropout = torch.nn.Dropout(pdrop)
    m.gemm_mode = _mode("bf16x3")
    tr = Trainer(m, dropout=True, lr=0.0)   # lr 0: the parameters stay put, only the masks change between replays
    batch = to_device(kat_batch(hp, 16, 40, seed=5), DEV)
    gs = tr.graphed(batch, warmup=2)
    a = float(gs.step(batch))
    b = float(gs.step(batch))
    assert a != b and abs(a - b) < 0.2 * abs(a)
