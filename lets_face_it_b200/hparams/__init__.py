"""Hyperparameter loading: the reference's yaml layout (`hparams/*.yaml`) into a Namespace (utils.py:13-41)."""
import argparse
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
FINAL_MODEL = os.path.join(HERE, "final_model.yaml")


def load_hparams(path=FINAL_MODEL):
    with open(path) as f:
        hp = yaml.safe_load(f)
    if not hp["Glow"].get("rnn_type"):
        hp["Glow"]["rnn_type"] = "gru"  # utils.py:32-33
    return argparse.Namespace(**hp)
