"""Debug helper: tensor-core backward flow core vs the FFMA backward pipeline on the same forward (bf16x3 mode)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import final_hparams
from tests.kat import build_kat_model, kat_batch, to_device
from lets_face_it_b200 import _cabi as cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hp = final_hparams()
m = build_kat_model(hp, "cuda:0"); m.glow.set_actnorm_init(True); m.gemm_mode = cabi.GEMM_BF16X3; m.train()
batch = to_device(kat_batch(hp, B, 80, seed=14), "cuda:0")
def run():
    m.zero_grad(); z, loss, _ = m(batch); loss.backward(); torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in m.named_parameters()}, float(loss)
os.environ["LFI_CORE_TC_BWD"] = "0"; g0, l0 = run()
os.environ["LFI_CORE_TC_BWD"] = "1"; g1, l1 = run()
print("loss", l0, l1)
worst = []
for k in g0:
    ref = g0[k].double(); err = float((g1[k].double() - ref).norm() / ref.norm().clamp_min(1e-30))
    worst.append((err, k))
worst.sort(reverse=True)
for e, k in worst[:12]: print("%.3e  %s" % (e, k))
print("max rel L2 err", worst[0][0], "nan:", any(torch.isnan(v).any().item() for v in g1.values()))
