"""Gradient error and speed of single-bf16 backward GEMMs (LFI_BWD_BF16=1) against the split-bf16 parity mode."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import final_hparams
from tests.kat import build_kat_model, kat_batch, to_device
from lets_face_it_b200 import _cabi as cabi
hp = final_hparams()
m = build_kat_model(hp, "cuda:0"); m.glow.set_actnorm_init(True); m.gemm_mode = cabi.GEMM_BF16X3; m.train()
batch = to_device(kat_batch(hp, 256, 80, seed=14), "cuda:0")
def run():
    m.zero_grad(); z, loss, _ = m(batch); loss.backward(); torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters()}
os.environ["LFI_BWD_BF16"] = "0"; g0 = run()
os.environ["LFI_BWD_BF16"] = "1"; g1 = run()
worst = sorted(((float((g1[k].double() - g0[k].double()).norm() / g0[k].double().norm().clamp_min(1e-30)), k) for k in g0), reverse=True)
for e, k in worst[:8]: print("%.3e  %s" % (e, k))
tot0 = torch.sqrt(sum((v.double() ** 2).sum() for v in g0.values())); totd = torch.sqrt(sum(((g1[k] - g0[k]).double() ** 2).sum() for k in g0))
print("global rel L2 error %.3e" % float(totd / tot0))
