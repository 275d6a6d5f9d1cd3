#!/usr/bin/env python
"""Profiling helper: a short autoregressive sampling call between cudaProfilerStart/Stop (use with ncu --profile-from-start off)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lets_face_it_b200 import _cabi as cabi  # noqa: E402
from lets_face_it_b200.hparams import load_hparams  # noqa: E402
from oracle import glow_oracle as O  # noqa: E402
from tests.kat import build_kat_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gemm", default="bf16x3")
ap.add_argument("--seqs", type=int, default=1024)
ap.add_argument("--frames", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hp = load_hparams()
hy = O.Hyper.from_hparams(hp)
m = build_kat_model(hp).to(dev).eval()
m.glow.set_actnorm_init(True)
m.gemm_mode = {"fp32": cabi.GEMM_FP32, "bf16x3": cabi.GEMM_BF16X3, "bf16": cabi.GEMM_BF16}[a.gemm]
T = hy.start_ts + a.frames
data = {k: v.to(dev) for k, v in O.synthetic_batch(hy, a.seqs, T, seed=5).items()}
data["p1_face"] = torch.zeros(a.seqs, hy.start_ts, hy.C, device=dev)
m.hparams.Infer["eps"] = 0.7
m.inference(T, data=data)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.inference(T, data=data)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
