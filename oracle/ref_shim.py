"""TEST INFRASTRUCTURE ONLY — import shim for the *live* reference (container only).

Makes `/root/reference/code/glow_pytorch/glow/{models,modules}.py` importable without editing
the reference: `glow/__init__.py:2` pulls in `glow/utils.py`, which imports `jsmin` and
`pytorch_lightning.Trainer` (utils.py:7,9); both are absent here, so two stub modules are
inserted first.  `/root/reference` does not exist on the GPU box: nothing
that runs there may import this file.  Used by `oracle/make_golden.py` and by the
container-only pinning tests.
"""
import argparse
import os
import sys
import types

REF_ROOT = os.environ.get("LFI_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")
REF_YAML_DIR = os.path.join(REF_CODE, "glow_pytorch", "hparams")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "glow_pytorch", "glow", "models.py"))


def import_reference():
    """Returns (models, modules) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "jsmin" not in sys.modules:
        m = types.ModuleType("jsmin")
        m.jsmin = lambda s: s
        sys.modules["jsmin"] = m
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class Trainer:  # only add_argparse_args is touched at import time
            @staticmethod
            def add_argparse_args(p):
                return p

        pl.Trainer = Trainer
        pl.LightningModule = object
        sys.modules["pytorch_lightning"] = pl
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    from glow_pytorch.glow import models, modules  # noqa: E402

    return models, modules


def load_reference_hparams(name="final_model.yaml"):
    import yaml

    with open(os.path.join(REF_YAML_DIR, name)) as f:
        return argparse.Namespace(**yaml.safe_load(f))
