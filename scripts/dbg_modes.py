import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import glow_oracle as O
from tests.helpers import final_hparams, relerr
from tests.kat import build_kat_model, kat_batch, oracle_params_from, to_device
from lets_face_it_b200 import _cabi as cabi
DEV = "cuda:0"
hp = final_hparams(); hy = O.Hyper.from_hparams(hp)
m = build_kat_model(hp); m.glow.set_actnorm_init(True)
P = O.clone_params(oracle_params_from(m), requires_grad=True)
B, T = 32, 40
batch = kat_batch(hp, B, T, seed=3)
z_ref, nll_ref, loss_ref = O.seq_forward(P, hy, batch); loss_ref.backward()
m = m.to(DEV).train()
for name, mode in (("fp32", cabi.GEMM_FP32), ("bf16x3", cabi.GEMM_BF16X3), ("bf16", cabi.GEMM_BF16)):
    m.gemm_mode = mode
    m.zero_grad()
    z_seq, loss, losses = m(to_device(batch, DEV)); loss.backward()
    print(name, "z", relerr(torch.stack(z_seq), z_ref.detach()), "nll", relerr(torch.stack(losses), nll_ref.detach()))
    errs = []
    for n, p in m.named_parameters():
        ref = P[n].grad
        errs.append((relerr(p.grad.reshape(ref.shape), ref), n, float(ref.abs().max())))
    errs.sort(reverse=True)
    for e in errs[:12]:
        print("   %.3e %s (max|g| %.3e)" % e)
