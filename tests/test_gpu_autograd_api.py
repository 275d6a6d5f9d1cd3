"""GPU: the module-level API under autograd (missing #3).  The reference's FlowStep / FlowNet / Glow.forward are
plain torch modules, so a per-frame training loop over them (models.py:305-342, 444-451; test_modules.py:30-67 drives exactly
this API) gets gradients; here every FlowStep.forward call is one autograd node backed by lfi_flowstep_fwd_train /
lfi_flowstep_bwd.  Checked against the oracle's flow_step differentiated by torch: gradients with respect to the input
frame, the conditioning vector, the carried RNN state (BPTT across frames) and every parameter."""
import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from tests.helpers import relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _hyper(C, K, H, D, rnn, coupling):
    return O.Hyper(C=C, K=K, H=H, D=D, rnn_type=rnn, scale_eps=1e-4, actnorm_scale=1.0, LU=True, coupling=coupling,
                   hist={m: 1 for m in O.MODALITIES}, enc={m: "none" for m in O.MODALITIES}, enc_hidden={m: 0 for m in O.MODALITIES},
                   in_dim={m: 1 for m in O.MODALITIES}, dropout={m: 0.0 for m in O.MODALITIES})


def _perturb(net):
    gen = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if "final_linear" in n or "actnorm" in n:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.15)


@pytest.mark.parametrize("rnn,coupling,C", [("gru", "affine", 56), ("lstm", "affine", 10), ("gru", "additive", 7)])
def test_flownet_per_frame_training_loop_gradients(rnn, coupling, C):
    """FlowNet (K = 3) driven frame by frame for three frames with the RNN state carried on the modules, loss = sum of a
    Gaussian NLL over the frames: d loss / d (inputs, conditioning, parameters) against torch autograd on the oracle."""
    from lets_face_it_b200.glow import FlowNet

    np.random.seed(3)
    torch.manual_seed(3)
    K, H, D, Fr, B, T = 3, 16, 24, 40, 21, 3
    net = FlowNet(C, H, D, K, 1, flow_permutation="invconv", flow_coupling=coupling, LU_decomposed=True, scale_eps=1e-4,
                  feature_encoder_dim=Fr, glow_rnn_type=rnn)
    _perturb(net)
    for l in net.layers:
        l.actnorm.inited = True
    hy = _hyper(C, K, H, D, rnn, coupling)
    P = O.clone_params({"glow.flow." + k: v for k, v in net.state_dict().items()}, requires_grad=True)
    xs = [torch.randn(B, C, requires_grad=True) for _ in range(T)]
    conds = [torch.randn(B, Fr, requires_grad=True) for _ in range(T)]
    # oracle
    state = {}
    loss_ref = 0.0
    for x, c in zip(xs, conds):
        z, ld = O.flow_encode(P, hy, x, c, state)
        loss_ref = loss_ref + O.nll_bits(ld, z).mean()
    loss_ref.backward()
    # device
    net = net.to(DEV).train()
    dx = [x.detach().to(DEV).requires_grad_(True) for x in xs]
    dc = [c.detach().to(DEV).requires_grad_(True) for c in conds]
    net.init_rnn_hidden()
    loss = 0.0
    for x, c in zip(dx, dc):
        z, ld = net(x, c, logdet=torch.zeros(B, device=DEV), reverse=False)
        logp = (-0.5 * (z ** 2 + float(np.log(2 * np.pi)))).sum(dim=1)
        loss = loss + (-(ld + logp) / float(np.log(2.0))).mean()
    assert abs(float(loss) - float(loss_ref)) < 1e-4 * abs(float(loss_ref))
    loss.backward()
    for a, b in zip(dx, xs):
        assert relerr(a.grad, b.grad) < 2e-3
    for a, b in zip(dc, conds):
        assert relerr(a.grad, b.grad) < 2e-3
    for n, p in net.named_parameters():
        ref = P["glow.flow." + n].grad
        assert p.grad is not None, n
        assert relerr(p.grad.reshape(ref.shape), ref) < 3e-3, n


def test_flowstep_forward_values_unchanged_under_autograd():
    """The autograd node returns the same y / logdet as the inference-style call (and as the oracle)."""
    from lets_face_it_b200.glow import FlowStep

    np.random.seed(4)
    torch.manual_seed(4)
    C, H, D, Fr, B = 56, 128, 64, 80, 33
    step = FlowStep(C, H, D, flow_permutation="invconv", flow_coupling="affine", LU_decomposed=True, scale_eps=1e-4,
                    feature_encoder_dim=Fr, glow_rnn_type="gru")
    _perturb(step)
    step.actnorm.inited = True
    step = step.to(DEV).train()
    x, c = torch.randn(B, C, device=DEV), torch.randn(B, Fr, device=DEV)
    with torch.no_grad():
        y0, ld0 = step(x, c, torch.zeros(B, device=DEV), False)
    step.init_rnn_hidden()
    y1, ld1 = step(x, c, torch.zeros(B, device=DEV), False)
    assert y1.requires_grad and ld1.requires_grad
    assert relerr(y1, y0) < 1e-6 and relerr(ld1, ld0) < 1e-6
