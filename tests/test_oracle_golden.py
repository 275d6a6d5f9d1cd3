"""CPU: pins oracle/glow_oracle.py to vectors produced by the unmodified reference
(oracle/make_golden.py), and — in the authoring container — to the live reference."""
import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from tests.helpers import golden_batch, golden_params, load_golden, relerr, small_hparams


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_oracle_forward_matches_reference_vectors(rnn):
    g = load_golden("kat_small_" + rnn)
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    P, batch = golden_params(g), golden_batch(g)
    with torch.no_grad():
        z, nll, loss = O.seq_forward(P, hy, batch)
    assert relerr(z, g["z"]) < 2e-6
    assert relerr(nll, g["nll"]) < 1e-6
    assert abs(float(loss) - float(g["loss"][0])) < 1e-4 * abs(float(g["loss"][0]))


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_oracle_sampling_and_invert_match_reference_vectors(rnn):
    g = load_golden("kat_small_" + rnn)
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    P, batch = golden_params(g), golden_batch(g)
    data = dict(batch)
    data["p1_face"] = torch.zeros(int(g["B"]), hy.start_ts, hy.C)
    L = int(g["infer_len"])
    x0 = O.seq_inference(P, hy, data, L, eps=0.0, noise=torch.zeros(L - hy.start_ts, int(g["B"]), hy.C))
    assert relerr(x0, g["x_eps0"]) < 5e-6
    x1 = O.seq_inference(P, hy, data, L, noise=torch.from_numpy(g["noise"]))
    assert relerr(x1, g["x_noise"]) < 5e-6
    rec, _ = O.seq_invert(P, hy, list(torch.from_numpy(g["z"])), batch)
    assert relerr(rec, g["rec"]) < 5e-5


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_oracle_backward_matches_reference_vectors(rnn):
    g = load_golden("kat_small_" + rnn)
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    P = O.clone_params(golden_params(g), requires_grad=True)
    batch = golden_batch(g)
    _, _, loss = O.seq_forward(P, hy, batch)
    loss.backward()
    for n in g["grad_names"]:
        ref = torch.from_numpy(g["grad/" + str(n)])
        got = P["" + str(n)].grad.reshape(ref.shape)
        assert relerr(got, ref) < 2e-4, n


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_oracle_dropout_masks_match_reference_vectors(rnn):
    g = load_golden("kat_small_" + rnn)
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    P = O.clone_params(golden_params(g), requires_grad=True)
    batch = golden_batch(g)
    masks = O.make_masks(hy, int(g["B"]), int(g["T"]) - hy.start_ts, seed=3)
    z, nll, loss = O.seq_forward(P, hy, batch, masks)
    loss.backward()
    assert relerr(z.detach(), g["masked_z"]) < 2e-6
    assert relerr(nll.detach(), g["masked_nll"]) < 1e-6
    for n in g["grad_names"]:
        ref = torch.from_numpy(g["masked_grad/" + str(n)])
        assert relerr(P[str(n)].grad.reshape(ref.shape), ref) < 2e-4, n


def test_oracle_ddi_matches_reference_vectors():
    """ActNorm data-dependent init (modules.py:32-43): start from the pre-DDI parameters."""
    g = load_golden("kat_small_gru")
    hy = O.Hyper.from_hparams(small_hparams("gru"))
    P, batch = golden_params(g), golden_batch(g)
    for k in list(P):
        if ".actnorm." in k:
            P[k] = torch.zeros_like(P[k])
    O.ddi_init(P, hy, batch)
    for k in P:
        if ".actnorm." in k:
            assert relerr(P[k], g["param/" + k]) < 1e-5, k
