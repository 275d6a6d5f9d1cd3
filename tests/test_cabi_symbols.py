"""CPU: the C-ABI library loads and exports every symbol include/lfi_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from tests.helpers import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "lfi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lfi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lets_face_it_b200 import _cabi

    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(names) == set(_cabi.SYMBOLS), set(names) ^ set(_cabi.SYMBOLS)


def test_binding_loads_and_reports_sizes():
    from lets_face_it_b200 import _cabi
    from lets_face_it_b200.engine import MODALITIES
    from lets_face_it_b200.hparams import load_hparams

    L = _cabi.lib()
    assert L.lfi_abi_version() == _cabi.ABI_VERSION
    hp = load_hparams()
    sh = _cabi.Shape()
    sh.C, sh.K, sh.H, sh.D, sh.G, sh.affine, sh.scale_eps = 56, 16, 128, 512, 3, 1, 1e-4
    for i, m in enumerate(MODALITIES):
        c = hp.Conditioning[m]
        sh.hist[i] = c["history"]
        sh.dim[i] = c.get("dim", hp.Data["speech_dim"])
        sh.ehid[i] = c["hidden_dim"] if c["enc"] == "rnn" else 0
    assert L.lfi_feature_dim(ctypes.byref(sh)) == 1560      # F
    assert L.lfi_feature_dim_folded(ctypes.byref(sh)) == 920  # F_eff
    assert L.lfi_start_ts(ctypes.byref(sh)) == 24
    assert L.lfi_coupling_out(ctypes.byref(sh)) == 56
    assert L.lfi_train_ws_bytes(ctypes.byref(sh), 256, 80, 0) > 0
    bad = _cabi.Shape()
    bad.C, bad.K, bad.H, bad.D, bad.G = 56, 16, 130, 512, 3   # H not a multiple of 4
    assert L.lfi_derived_bytes(ctypes.byref(bad)) == 0
    assert b"H=130" in L.lfi_last_error()


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the module API raises instead of silently computing on the host."""
    import pytest
    import torch

    from lets_face_it_b200.glow import ActNorm2d, SeqGlow
    from lets_face_it_b200.hparams import load_hparams
    from tests.helpers import small_hparams

    with pytest.raises(RuntimeError):
        ActNorm2d(8)(torch.zeros(2, 8), 0)
    m = SeqGlow(small_hparams())
    batch = {"p1_face": torch.zeros(2, 12, 12), "p2_face": torch.zeros(2, 12, 12), "p1_speech": torch.zeros(2, 12, 5),
             "p2_speech": torch.zeros(2, 12, 5)}
    with pytest.raises(RuntimeError):
        m(batch)


def test_dropin_constructor_reproduces_reference_parameters():
    """Same seeds => bit-identical parameters and state-dict keys as the reference (golden fingerprints)."""
    import numpy as np

    from tests.helpers import final_hparams, load_golden
    from tests.kat import build_kat_model, fingerprint

    g = load_golden("kat_full")
    m = build_kat_model(final_hparams())
    names, fp = fingerprint(m.state_dict())
    assert names == [str(x) for x in g["fp_names"]]
    assert np.abs(fp - g["fp_init"]).max() == 0.0
