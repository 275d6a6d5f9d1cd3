"""Shared helpers for the parity tests (golden loading, hparams for the golden cases)."""
import argparse
import copy
import os

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
YAML = os.path.join(ROOT, "lets_face_it_b200", "hparams", "final_model.yaml")


def final_hparams():
    with open(YAML) as f:
        return argparse.Namespace(**yaml.safe_load(f))


def small_hparams(rnn_type="gru"):
    """Must match oracle/make_golden.py:small_hparams."""
    hp = copy.deepcopy(final_hparams())
    c = hp.Conditioning
    c["cond_dim"] = 24
    c["p1_face"].update(dim=12, history=3)
    c["p2_face"].update(dim=12, history=6, hidden_dim=8)
    c["p1_speech"].update(history=2, hidden_dim=6)
    c["p2_speech"].update(history=4, hidden_dim=8)
    hp.Data["speech_dim"] = 5
    hp.Glow.update(K=3, hidden_channels=16, rnn_type=rnn_type)
    hp.Validation["scale_logging"] = False
    return hp


def load_golden(tag):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)
    return g


def golden_params(g):
    return {k[len("param/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param/")}


def golden_batch(g):
    return {k[len("batch/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("batch/")}


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
