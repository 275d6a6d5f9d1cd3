"""GPU parity: the CUDA path (through the module API -> C ABI) against the reference-generated golden vectors
and against the CPU oracle on the same seeded inputs.  Tolerances (fp32 mode, BASELINE.json north star):
z / x / NLL within 1e-4 relative (relative to the tensor's max magnitude); gradients within 2e-3 relative per
tensor (the reference itself moves by ~1e-6 between thread counts)."""
import copy
import ctypes

import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from tests.helpers import final_hparams, golden_batch, golden_params, load_golden, relerr, small_hparams
from tests.kat import build_kat_model, kat_batch, oracle_params_from, to_device

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _model_from_golden(rnn):
    from lets_face_it_b200.glow import ModalityEncoder, SeqGlow

    g = load_golden("kat_small_" + rnn)
    m = SeqGlow(small_hparams(rnn))
    m.load_state_dict(golden_params(g))
    for mod in m.modules():
        if isinstance(mod, ModalityEncoder):
            mod.dropout = None
    m.glow.set_actnorm_init(True)
    return m.to(DEV), g


# ------------------------------------------------------------------------------------------------ primitives
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_fp32_tiles_match_torch(tA, tB):
    from lets_face_it_b200 import _cabi as cabi

    torch.manual_seed(0)
    L = cabi.lib()
    for (M, N, K, batch) in [(130, 70, 33, 1), (257, 384, 540, 2), (64, 56, 1000, 3), (5, 3, 2, 1)]:
        A = torch.randn(batch, *((K, M) if tA else (M, K)), device=DEV)
        B = torch.randn(batch, *((N, K) if tB else (K, N)), device=DEV)
        bias = torch.randn(batch, N, device=DEV)
        aux = torch.randn(batch, M, N, device=DEV)
        C = torch.zeros(batch, M, N, device=DEV)
        opA = A.transpose(1, 2) if tA else A
        opB = B.transpose(1, 2) if tB else B
        ref = torch.matmul(opA.double(), opB.double())
        for epi in (0, cabi.EPI_BIAS, cabi.EPI_BIAS | cabi.EPI_LRELU, cabi.EPI_LRELU_BWD, cabi.EPI_ACCUM):
            C.fill_(0.5)
            want = ref.clone()
            if epi & cabi.EPI_BIAS:
                want = want + bias[:, None, :].double()
            if epi & cabi.EPI_LRELU:
                want = torch.where(want > 0, want, 0.01 * want)
            if epi & cabi.EPI_LRELU_BWD:
                want = want * torch.where(aux > 0, 1.0, 0.01).double()
            if epi & cabi.EPI_ACCUM:
                want = want + 0.5
            cabi.check(L.lfi_gemm(cabi.GEMM_FP32, tA, tB, M, N, K, A.data_ptr(), A.shape[2], A[0].numel(), B.data_ptr(), B.shape[2],
                                  B[0].numel(), C.data_ptr(), N, M * N, bias.data_ptr(), N, aux.data_ptr(), N, M * N, batch, epi, None, 0,
                                  cabi.stream_ptr()), "lfi_gemm")
            assert relerr(C, want) < 2e-6, (M, N, K, batch, epi)


def test_gemm_splitk_accumulate():
    from lets_face_it_b200 import _cabi as cabi

    torch.manual_seed(1)
    M, N, K = 96, 40, 5000
    A = torch.randn(K, M, device=DEV)
    B = torch.randn(K, N, device=DEV)
    C = torch.full((M, N), 2.0, device=DEV)
    cabi.check(cabi.lib().lfi_gemm(cabi.GEMM_FP32, 1, 0, M, N, K, A.data_ptr(), M, 0, B.data_ptr(), N, 0, C.data_ptr(), N, 0, None, 0, None, 0,
                                   0, 1, cabi.EPI_ACCUM, None, 0, cabi.stream_ptr()), "lfi_gemm")
    assert relerr(C, A.double().t() @ B.double() + 2.0) < 5e-6


def test_invconv_compose_and_chain_rule_match_oracle():
    from lets_face_it_b200 import _cabi as cabi
    from lets_face_it_b200.glow import InvertibleConv1x1

    np.random.seed(5)
    K, C = 3, 12
    convs = [InvertibleConv1x1(C, LU_decomposed=True) for _ in range(K)]
    P = {}
    for k, c in enumerate(convs):
        for n in ("l", "u", "log_s", "p", "sign_s"):
            t = getattr(c, n).detach().clone()
            if n in ("l", "u"):
                t = t + 0.05 * torch.randn_like(t)
            P["%d.%s" % (k, n)] = t
    st = lambda n: torch.stack([P["%d.%s" % (k, n)] for k in range(K)]).to(DEV).contiguous()
    p, l, u, ls, sg = st("p"), st("l"), st("u"), st("log_s"), st("sign_s")
    W = torch.empty(K, C, C, device=DEV)
    Wi = torch.empty(K, C, C, device=DEV)
    L = cabi.lib()
    ws = torch.empty(L.lfi_invconv_ws_bytes(K, C), dtype=torch.uint8, device=DEV)
    cabi.check(L.lfi_invconv_compose(K, C, p.data_ptr(), l.data_ptr(), u.data_ptr(), ls.data_ptr(), sg.data_ptr(), W.data_ptr(), Wi.data_ptr(),
                                     ws.data_ptr(), ws.numel(), cabi.stream_ptr()), "compose")
    dW = torch.randn(K, C, C, device=DEV)
    dl, du, dls = torch.zeros_like(l), torch.zeros_like(u), torch.zeros_like(ls)
    cabi.check(L.lfi_invconv_compose_bwd(K, C, p.data_ptr(), l.data_ptr(), u.data_ptr(), ls.data_ptr(), sg.data_ptr(), dW.data_ptr(),
                                         dl.data_ptr(), du.data_ptr(), dls.data_ptr(), ws.data_ptr(), ws.numel(), cabi.stream_ptr()), "compose_bwd")
    for k in range(K):
        Pk = {"x." + n: P["%d.%s" % (k, n)].clone().requires_grad_(n in ("l", "u", "log_s")) for n in ("l", "u", "log_s", "p", "sign_s")}
        w, _ = O.invconv_weight(Pk, "x.", C, False)
        wi, _ = O.invconv_weight(Pk, "x.", C, True)
        assert relerr(W[k], w.detach()) < 1e-6
        assert relerr(Wi[k], wi.detach()) < 1e-5
        (w * dW[k].cpu()).sum().backward()
        assert relerr(dl[k], Pk["x.l"].grad) < 1e-5
        assert relerr(du[k], Pk["x.u"].grad) < 1e-5
        assert relerr(dls[k], Pk["x.log_s"].grad) < 1e-5


def test_clip_adam_matches_torch():
    from lets_face_it_b200 import _cabi as cabi

    torch.manual_seed(3)
    n = 100003
    theta = torch.randn(n, device=DEV)
    ref = theta.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.9999), eps=1e-8)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    scratch = torch.zeros(2, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV) * 3.0
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 20.0)
        opt.step()
        cabi.check(cabi.lib().lfi_clip_adam(theta.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.9999, 1e-8, 20.0, 1.0,
                                            step, scratch.data_ptr(), cabi.stream_ptr()), "clip_adam")
        assert relerr(theta, ref.detach()) < 2e-6, step


# ------------------------------------------------------------------------------------------------ module API (test_modules.py of the reference)
def test_module_roundtrips_like_reference_test_modules():
    from lets_face_it_b200.glow import ActNorm2d, FlowNet, FlowStep, InvertibleConv1x1

    np.random.seed(0)
    torch.manual_seed(0)
    actnorm = ActNorm2d(54).to(DEV)
    x = torch.Tensor(np.random.rand(2, 54)).to(DEV)
    actnorm.initialize_parameters(x)
    y, det = actnorm(x, 0)
    x_, _ = actnorm(y, None, True)
    assert float((x_ - x).abs().max()) < 1e-5

    conv = InvertibleConv1x1(96).to(DEV)
    x = torch.Tensor(np.random.rand(2, 96)).to(DEV)
    y, det = conv(x, 0)
    x_, _ = conv(y, None, True)
    assert float((x_ - x).abs().max()) < 1e-5

    for lu in (False, True):
        step = FlowStep(54, 256, flow_permutation="invconv", flow_coupling="affine", cond_dim=32, feature_encoder_dim=64,
                        glow_rnn_type="gru", LU_decomposed=lu).to(DEV)
        with torch.no_grad():
            step.f.final_linear.weight.normal_(0, 0.05)
            step.f.final_linear.bias.normal_(0, 0.05)
        x = torch.Tensor(np.random.rand(2, 54)).to(DEV)
        cond = torch.Tensor(np.random.rand(2, 64)).to(DEV)
        y, det = step(x, cond, 0, False)
        step.init_rnn_hidden()
        x_, det0 = step(y, cond, det, True)
        assert float((x_ - x).abs().max()) < 2e-5
        assert float(det0.abs().max()) < 1e-3  # forward + reverse log-dets cancel

    net = FlowNet(C=54, hidden_channels=256, cond_dim=64, K=3, L=1, feature_encoder_dim=32, glow_rnn_type="gru").to(DEV)
    x = torch.Tensor(np.random.rand(4, 54)).to(DEV)
    cond = torch.Tensor(np.random.rand(4, 32)).to(DEV)
    y, det = net(x, cond)
    net.init_rnn_hidden()
    x_, det0 = net(y, cond, reverse=True)
    assert float((x_ - x).abs().max()) < 2e-5


@pytest.mark.parametrize("rnn,coupling,C", [("gru", "affine", 12), ("lstm", "affine", 13), ("gru", "additive", 12)])
def test_flowstep_matches_oracle(rnn, coupling, C):
    """Single FlowStep, two consecutive frames (carried RNN state), forward and reverse, vs the CPU oracle."""
    from lets_face_it_b200.glow import FlowStep

    np.random.seed(2)
    torch.manual_seed(2)
    H, D, Fr, B = 16, 24, 40, 37
    step = FlowStep(C, H, D, flow_permutation="invconv", flow_coupling=coupling, LU_decomposed=True, scale_eps=1e-4,
                    feature_encoder_dim=Fr, glow_rnn_type=rnn)
    with torch.no_grad():
        step.f.final_linear.weight.normal_(0, 0.1)
        step.f.final_linear.bias.normal_(0, 0.1)
        step.f.final_linear.logs.normal_(0, 0.1)
        step.actnorm.bias.normal_(0, 0.2)
        step.actnorm.logs.normal_(0, 0.2)
    step.actnorm.inited = True
    P = {"glow.flow.layers.0." + k: v.detach().clone() for k, v in step.state_dict().items()}
    hy = O.Hyper(C=C, K=1, H=H, D=D, rnn_type=rnn, scale_eps=1e-4, actnorm_scale=1.0, LU=True, coupling=coupling,
                 hist={m: 1 for m in O.MODALITIES}, enc={m: "none" for m in O.MODALITIES}, enc_hidden={m: 0 for m in O.MODALITIES},
                 in_dim={m: 1 for m in O.MODALITIES}, dropout={m: 0.0 for m in O.MODALITIES})
    step = step.to(DEV).eval()
    xs = [torch.randn(B, C) for _ in range(2)]
    conds = [torch.randn(B, Fr) for _ in range(2)]
    state = {}
    outs = []
    for x, c in zip(xs, conds):
        with torch.no_grad():
            y_ref, ld_ref = O.flow_step(P, hy, 0, x, c, torch.zeros(B), False, state)
        y, ld = step(x.to(DEV), c.to(DEV), torch.zeros(B, device=DEV), False)
        assert relerr(y, y_ref) < TOL and relerr(ld, ld_ref) < TOL
        outs.append(y_ref)
    step.init_rnn_hidden()
    state = {}
    for y_ref, x, c in zip(outs, xs, conds):
        with torch.no_grad():
            x_ref, ld_ref = O.flow_step(P, hy, 0, y_ref, c, torch.zeros(B), True, state)
        x_, ld = step(y_ref.to(DEV), c.to(DEV), torch.zeros(B, device=DEV), True)
        assert relerr(x_, x_ref) < TOL and relerr(ld, ld_ref) < TOL
        assert relerr(x_, x) < 1e-3


# ------------------------------------------------------------------------------------------------ sequence paths vs golden (small)
@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_seq_forward_matches_reference_vectors(rnn):
    m, g = _model_from_golden(rnn)
    m.eval()
    batch = to_device(golden_batch(g), DEV)
    with torch.no_grad():
        z_seq, loss, losses = m(batch)
    z = torch.stack(z_seq)
    assert len(z_seq) == g["z"].shape[0] and losses[0].device.type == "cpu" and loss.shape == (1,)
    assert relerr(z, g["z"]) < TOL
    assert relerr(torch.stack(losses), g["nll"]) < TOL
    assert abs(float(loss) - float(g["loss"][0])) < TOL * abs(float(g["loss"][0]))


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_seq_sampling_and_invert_match_reference_vectors(rnn):
    m, g = _model_from_golden(rnn)
    m.eval()
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    batch = to_device(golden_batch(g), DEV)
    data = dict(batch)
    B, L = int(g["B"]), int(g["infer_len"])
    data["p1_face"] = torch.zeros(B, hy.start_ts, hy.C, device=DEV)
    m.hparams.Infer["eps"] = 0
    x0 = m.inference(L, data=data)
    assert tuple(x0.shape) == g["x_eps0"].shape
    assert relerr(x0, g["x_eps0"]) < TOL
    x1 = m.inference(L, data=data, noise=torch.from_numpy(g["noise"]).to(DEV))
    assert relerr(x1, g["x_noise"]) < TOL
    rec, bl = m.invert([t.to(DEV) for t in torch.from_numpy(g["z"])], batch)
    assert relerr(torch.stack(rec), g["rec"]) < 5e-4
    assert abs(float(bl) - float(g["invert_loss"][0])) < TOL * abs(float(g["invert_loss"][0]))
    # round trip: invert(forward(x)) == x
    assert relerr(torch.stack(rec), batch["p1_face"][:, hy.start_ts:].transpose(0, 1)) < 5e-4


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_seq_backward_matches_reference_vectors(rnn):
    m, g = _model_from_golden(rnn)
    m.train()
    batch = to_device(golden_batch(g), DEV)
    m.zero_grad()
    _, loss, _ = m(batch)
    loss.backward()
    assert abs(float(loss) - float(g["loss_train"][0])) < TOL * abs(float(g["loss_train"][0]))
    named = dict(m.named_parameters())
    worst = 0.0
    for n in g["grad_names"]:
        n = str(n)
        ref = torch.from_numpy(g["grad/" + n])
        assert named[n].grad is not None, n
        e = relerr(named[n].grad.reshape(ref.shape), ref)
        worst = max(worst, e)
        assert e < 2e-3, (n, e)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))
    assert abs(float(total) - float(g["grad_total_norm"])) < 1e-3 * float(g["grad_total_norm"])


@pytest.mark.parametrize("rnn", ["gru", "lstm"])
def test_seq_dropout_masks_forward_backward(rnn):
    m, g = _model_from_golden(rnn)
    m.train()
    hy = O.Hyper.from_hparams(small_hparams(rnn))
    batch = to_device(golden_batch(g), DEV)
    masks = O.make_masks(hy, int(g["B"]), int(g["T"]) - hy.start_ts, seed=3)
    m.injected_masks = {k: (v.to(DEV) if v is not None else None) for k, v in masks.items()}
    m.zero_grad()
    z_seq, loss, losses = m(batch)
    loss.backward()
    assert relerr(torch.stack(z_seq), g["masked_z"]) < TOL
    assert relerr(torch.stack(losses), g["masked_nll"]) < TOL
    named = dict(m.named_parameters())
    for n in g["grad_names"]:
        n = str(n)
        ref = torch.from_numpy(g["masked_grad/" + n])
        assert relerr(named[n].grad.reshape(ref.shape), ref) < 2e-3, n


def test_actnorm_ddi_matches_reference_vectors():
    from lets_face_it_b200.glow import ModalityEncoder, SeqGlow

    g = load_golden("kat_small_gru")
    m = SeqGlow(small_hparams("gru"))
    sd = golden_params(g)
    for k in sd:
        if ".actnorm." in k:
            sd[k] = torch.zeros_like(sd[k])
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, ModalityEncoder):
            mod.dropout = None
    m = m.to(DEV).train()
    batch = to_device(golden_batch(g), DEV)
    _, loss, _ = m(batch)
    assert all(l.actnorm.inited for l in m.glow.flow.layers)
    for k, v in m.state_dict().items():
        if ".actnorm." in k:
            assert relerr(v, g["param/" + k]) < 1e-4, k
    assert abs(float(loss) - float(g["loss_ddi"][0])) < TOL * abs(float(g["loss_ddi"][0]))


# ------------------------------------------------------------------------------------------------ shape sweep vs oracle
@pytest.mark.parametrize("cfg", [
    dict(C=10, K=2, H=8, D=12, rnn="gru", coupling="affine", B=3, T=9),
    dict(C=7, K=3, H=20, D=16, rnn="lstm", coupling="affine", B=33, T=10),
    dict(C=16, K=2, H=132, D=40, rnn="gru", coupling="additive", B=70, T=9),
    dict(C=56, K=4, H=128, D=64, rnn="gru", coupling="affine", B=17, T=10),
])
def test_seq_paths_match_oracle_over_shapes(cfg):
    """Edge shapes: odd C, ragged tiles (B not a multiple of the tile), additive coupling, H > one column chunk."""
    from lets_face_it_b200.glow import ModalityEncoder, SeqGlow

    hp = copy.deepcopy(small_hparams(cfg["rnn"]))
    hp.Conditioning["p1_face"]["dim"] = cfg["C"]
    hp.Conditioning["cond_dim"] = cfg["D"]
    hp.Glow.update(K=cfg["K"], hidden_channels=cfg["H"], flow_coupling=cfg["coupling"])
    torch.manual_seed(11)
    np.random.seed(11)
    m = SeqGlow(hp)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    for mod in m.modules():
        if isinstance(mod, ModalityEncoder):
            mod.dropout = None
    m.glow.set_actnorm_init(True)
    hy = O.Hyper.from_hparams(hp)
    P = O.clone_params(oracle_params_from(m), requires_grad=True)
    batch = O.synthetic_batch(hy, cfg["B"], cfg["T"], seed=4)
    z_ref, nll_ref, loss_ref = O.seq_forward(P, hy, batch)
    loss_ref.backward()
    m = m.to(DEV).train()
    z_seq, loss, losses = m(to_device(batch, DEV))
    loss.backward()
    assert relerr(torch.stack(z_seq), z_ref.detach()) < TOL
    assert relerr(torch.stack(losses), nll_ref.detach()) < TOL
    for n, p in m.named_parameters():
        ref = P[n].grad
        assert relerr(p.grad.reshape(ref.shape), ref) < 3e-3, n
    # sampling with noise + invert
    m.eval()
    Tg = cfg["T"] - hy.start_ts
    noise = torch.randn(Tg, cfg["B"], cfg["C"], generator=gen) * 0.7
    data = dict(batch)
    data["p1_face"] = batch["p1_face"][:, :hy.start_ts]
    with torch.no_grad():
        P0 = {k: v.detach() for k, v in P.items()}
        x_ref = O.seq_inference(P0, hy, data, cfg["T"], noise=noise)
    x = m.inference(cfg["T"], data=to_device(data, DEV), noise=noise.to(DEV))
    assert relerr(x, x_ref) < 5e-4


# ------------------------------------------------------------------------------------------------ full final_model.yaml KAT
def test_full_kat_final_model():
    """final_model.yaml, B=64, T=80: DDI pass, eval forward, sampling, backward vs the reference."""
    g = load_golden("kat_full")
    hp = final_hparams()
    m = build_kat_model(hp, DEV)
    B, T = int(g["B"]), int(g["T"])
    batch = to_device(kat_batch(hp, B, T), DEV)
    m.train()
    _, loss_ddi, _ = m(batch)
    assert abs(float(loss_ddi) - float(g["loss_ddi"][0])) < TOL * abs(float(g["loss_ddi"][0]))
    for k, v in m.state_dict().items():
        if ".actnorm." in k:
            assert relerr(v, g["param/" + k]) < 1e-4, k
    m.eval()
    with torch.no_grad():
        z_seq, loss, losses = m(batch)
    z = torch.stack(z_seq)
    assert abs(float(loss) - float(g["loss"][0])) < TOL * abs(float(g["loss"][0]))
    assert relerr(z[:, :8], g["z_head"]) < TOL
    assert relerr(torch.stack(losses), g["nll"]) < TOL
    assert abs(float(z.double().sum()) - float(g["z_sum"])) < 0.5
    # sampling, eps = 0 and injected noise
    data = dict(batch)
    data["p1_face"] = torch.zeros(B, 24, 56, device=DEV)
    m.hparams.Infer["eps"] = 0
    x0 = m.inference(int(g["infer_len"]), data=data)
    assert relerr(x0[:8], g["x_eps0_head"]) < TOL
    gn = torch.Generator().manual_seed(11)
    noise = torch.randn(int(g["infer_len"]) - 24, B, 56, generator=gn) * 0.7
    x1 = m.inference(int(g["infer_len"]), data=data, noise=noise.to(DEV))
    assert relerr(x1[:8], g["x_noise_head"]) < TOL
    # invert round trip
    rec, _ = m.invert(z_seq, batch)
    assert relerr(torch.stack(rec), batch["p1_face"][:, 24:].transpose(0, 1)) < 5e-4
    # backward: per-tensor gradient norms of the reference
    m.train()
    m.zero_grad()
    loss_t = m(batch)[1]
    loss_t.backward()
    named = dict(m.named_parameters())
    for n, ref in zip(g["grad_names"], g["grad_norms"]):
        got = float(named[str(n)].grad.double().norm())
        assert abs(got - ref) < 2e-3 * max(ref, 1e-3), (str(n), got, ref)
    total = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())))
    assert abs(total - float(g["grad_total_norm"])) < 1e-3 * float(g["grad_total_norm"])
