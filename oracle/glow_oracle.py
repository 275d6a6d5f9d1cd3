"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's conditional-Glow hot path.

This file is the *oracle* (checker) for the CUDA path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
it; the product package `lets_face_it_b200` never does.

It restates, as pure functions over a flat `{name: tensor}` parameter dict (the reference's
state-dict names), the algorithm of

    /root/reference/code/glow_pytorch/glow/modules.py   (ActNorm2d 10-80, LinearZeros 83-95,
                                                         InvertibleConv1x1 122-194, GaussianDiag 197-235)
    /root/reference/code/glow_pytorch/glow/models.py    (ModalityEncoder 12-80, FeatureEncoder 83-145,
                                                         f_seq 148-214, FlowStep 217-376, FlowNet 379-467,
                                                         Glow 470-521, SeqGlow 524-645)
    /root/reference/code/glow_pytorch/glow/thops.py     (split_feature 36-44)

frame by frame, in the reference's operation order, in fp32 on the CPU (torch ATen CPU kernels —
the same third-party arithmetic the reference calls).  The GRU/LSTM gate equations are PyTorch's
documented ones (`torch.gru_cell` / `torch.lstm_cell`; reference pins torch==1.6.0,
code/glow_pytorch/environment.yml:73; here torch 2.11).

Pinning: `oracle/make_golden.py` runs the unmodified reference in the authoring container and
writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file against those vectors
(and, when `/root/reference` is present, against the live reference).  Parity is therefore pinned
by reference outputs generated here, not by reference-owned golden vectors (the reference has
none: SURVEY.md §4).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

LN2 = math.log(2.0)
LOG2PI = math.log(2.0 * math.pi)
MODALITIES = ("p1_face", "p2_face", "p1_speech", "p2_speech")  # concat order, models.py:127-145


@dataclass
class Hyper:
    """Shape/config facts read from the hparams Namespace (final_model.yaml layout)."""

    C: int
    K: int
    H: int
    D: int
    rnn_type: str
    scale_eps: float
    actnorm_scale: float
    LU: bool
    coupling: str
    hist: Dict[str, int]
    enc: Dict[str, str]
    enc_hidden: Dict[str, int]
    in_dim: Dict[str, int]
    dropout: Dict[str, float]
    use_frame_nb: bool = False
    start_ts: int = field(init=False)
    F: int = field(init=False)

    def __post_init__(self):
        self.start_ts = max(self.hist.values())  # utils.py:44-50
        f = 0
        for m in MODALITIES:
            if m != "p1_face" and not self.hist[m]:
                continue
            f += self.enc_dim(m)
        if self.use_frame_nb:
            f += 1
        self.F = f

    def enc_dim(self, m):  # models.py:28,35,41,51
        e = self.enc[m]
        if e in ("rnn", "lstm"):
            return 2 * self.enc_hidden[m]
        if e == "none":
            return self.in_dim[m] * self.hist[m]
        raise NotImplementedError(e)

    @staticmethod
    def from_hparams(hp) -> "Hyper":
        cond, glow, data = hp.Conditioning, hp.Glow, hp.Data
        in_dim = {
            "p1_face": cond["p1_face"]["dim"],
            "p2_face": cond["p2_face"]["dim"],
            "p1_speech": data["speech_dim"],
            "p2_speech": data["speech_dim"],
        }
        return Hyper(
            C=cond["p1_face"]["dim"],
            K=glow["K"] * glow["L"],
            H=glow["hidden_channels"],
            D=cond["cond_dim"],
            rnn_type=glow.get("rnn_type") or "gru",
            scale_eps=float(glow["scale_eps"]),
            actnorm_scale=float(glow["actnorm_scale"]),
            LU=bool(glow["LU_decomposed"]),
            coupling=glow["flow_coupling"],
            hist={m: cond[m]["history"] for m in MODALITIES},
            enc={m: cond[m]["enc"] for m in MODALITIES},
            enc_hidden={m: cond[m]["hidden_dim"] for m in MODALITIES},
            in_dim=in_dim,
            dropout={m: float(cond[m]["dropout"]) for m in MODALITIES},
            use_frame_nb=bool(cond["use_frame_nb"]),
        )


# ----------------------------------------------------------------------------- primitives


def actnorm(P, pre, x, logdet, reverse):
    """modules.py:45-80.  logdet term is multiplied by x.size(1) (= C, the channel dim)."""
    bias, logs = P[pre + "bias"], P[pre + "logs"]
    if not reverse:
        x = (x + bias) * torch.exp(logs)
    else:
        x = x * torch.exp(-logs) - bias
    if logdet is not None:
        d = torch.sum(logs) * x.size(1)
        logdet = logdet - d if reverse else logdet + d
    return x, logdet


def actnorm_ddi(P, pre, x, scale):
    """modules.py:32-43 data-dependent init (training, first call)."""
    bias = -x.mean(dim=0, keepdim=True)
    var = ((x + bias) ** 2).mean(dim=0, keepdim=True)
    logs = torch.log(scale / (torch.sqrt(var) + 1e-6))
    for name, val in (("bias", bias), ("logs", logs)):  # `.data.copy_` in the reference: the parameter stays a trainable leaf
        old = P.get(pre + name)
        P[pre + name] = val.detach().clone().requires_grad_(bool(old is not None and old.requires_grad))


def invconv_weight(P, pre, C, reverse, LU=True):
    """modules.py:149-178: W = P (L*mask + I) (U*mask^T + diag(sign_s e^{log_s})); reverse uses the
    fp64 inverses of L and U (cast back to fp32) and P^-1."""
    if not LU:
        w = P[pre + "weight"]
        dlogdet = torch.slogdet(w)[1] * C
        if reverse:
            w = torch.inverse(w.double()).float()
        return w, dlogdet
    l_mask = torch.tril(torch.ones(C, C), -1)
    eye = torch.eye(C)
    l = P[pre + "l"] * l_mask + eye
    u = P[pre + "u"] * l_mask.t().contiguous() + torch.diag(P[pre + "sign_s"] * torch.exp(P[pre + "log_s"]))
    dlogdet = torch.sum(P[pre + "log_s"]) * C
    if not reverse:
        w = P[pre + "p"] @ (l @ u)
    else:
        li = torch.inverse(l.double()).float()
        ui = torch.inverse(u.double()).float()
        w = ui @ (li @ torch.inverse(P[pre + "p"]))
    return w, dlogdet


def invconv(P, pre, x, logdet, reverse, LU=True):
    """modules.py:180-194 (row-vector convention z = x @ W)."""
    w, d = invconv_weight(P, pre, x.size(1), reverse, LU)
    z = x @ w
    if logdet is not None:
        logdet = logdet - d if reverse else logdet + d
    return z, logdet


def linear_zeros(P, pre, h, logscale_factor=3):
    """modules.py:93-95."""
    return F.linear(h, P[pre + "weight"], P[pre + "bias"]) * torch.exp(P[pre + "logs"] * logscale_factor)


def coupling_net(P, pre, hy: Hyper, z1, cond, state, k):
    """f_seq.forward models.py:204-214.  `state[k]` is h (GRU) or (h, c) (LSTM); missing = zeros
    (the LSTM branch of the reference crashes on (None, None); zero state is its evident intent,
    SURVEY.md §0.1)."""
    c = F.leaky_relu(F.linear(cond, P[pre + "cond_transform.0.weight"], P[pre + "cond_transform.0.bias"]), 0.01)
    x = torch.cat((z1, c), dim=1)
    B = x.size(0)
    w_ih, w_hh = P[pre + "rnn.weight_ih"], P[pre + "rnn.weight_hh"]
    b_ih, b_hh = P[pre + "rnn.bias_ih"], P[pre + "rnn.bias_hh"]
    if hy.rnn_type == "gru":
        h0 = state.get(k)
        if h0 is None:
            h0 = x.new_zeros(B, hy.H)
        h = torch.gru_cell(x, h0, w_ih, w_hh, b_ih, b_hh)
        state[k] = h
    else:
        hc = state.get(k)
        if hc is None:
            hc = (x.new_zeros(B, hy.H), x.new_zeros(B, hy.H))
        h, cc = torch.lstm_cell(x, hc, w_ih, w_hh, b_ih, b_hh)
        state[k] = (h, cc)
    return linear_zeros(P, pre + "final_linear.", h)


def flow_step(P, hy: Hyper, k, x, cond, logdet, reverse, state, scales=None):
    """FlowStep.normal_flow / reverse_flow, models.py:311-373 (invconv permutation, affine or
    additive coupling)."""
    pre = "glow.flow.layers.%d." % k
    half = hy.C // 2
    if not reverse:
        z, logdet = actnorm(P, pre + "actnorm.", x, logdet, False)
        z, logdet = invconv(P, pre + "invconv.", z.float(), logdet, False, hy.LU)
        z1, z2 = z[:, :half], z[:, half:]
        h = coupling_net(P, pre + "f.", hy, z1, cond, state, k)
        if hy.coupling == "additive":
            z2 = z2 + h
        else:
            shift, scale = h[:, 0::2], h[:, 1::2]
            scale = torch.sigmoid(scale + 2.0).clamp(hy.scale_eps)
            if scales is not None:
                scales[k] = scale
            z2 = (z2 + shift) * scale
            logdet = torch.sum(torch.log(scale), dim=1) + logdet
        return torch.cat((z1, z2), dim=1), logdet
    z1, z2 = x[:, :half], x[:, half:]
    h = coupling_net(P, pre + "f.", hy, z1, cond, state, k)
    if hy.coupling == "additive":
        z2 = z2 - h
    else:
        shift, scale = h[:, 0::2], h[:, 1::2]
        scale = torch.sigmoid(scale + 2.0).clamp(hy.scale_eps)
        z2 = z2 / scale - shift
        logdet = -torch.sum(torch.log(scale), dim=1) + logdet
    z = torch.cat((z1, z2), dim=1)
    z, logdet = invconv(P, pre + "invconv.", z, logdet, True, hy.LU)
    z, logdet = actnorm(P, pre + "actnorm.", z, logdet, True)
    return z, logdet


def flow_encode(P, hy, x, cond, state, scales=None):
    """Glow.normal_flow + FlowNet.encode, models.py:449-451, 504-506."""
    logdet = torch.zeros_like(x[:, 0])
    for k in range(hy.K):
        x, logdet = flow_step(P, hy, k, x, cond, logdet, False, state, scales)
    return x, logdet


def flow_decode(P, hy, z, cond, state):
    """Glow.reverse_flow + FlowNet.decode, models.py:460-461, 508-513 (logdet starts at 0.0)."""
    logdet = 0.0
    for k in reversed(range(hy.K)):
        z, logdet = flow_step(P, hy, k, z, cond, logdet, True, state)
    return z, logdet


def nll_bits(logdet, z):
    """GaussianDiag.logp_simplified (modules.py:200-212) + SeqGlow.loss (models.py:563-565)."""
    obj = logdet + torch.sum(-0.5 * (z ** 2 + LOG2PI), dim=1)
    return (-obj) / LN2


# ----------------------------------------------------------------------------- conditioning


def encode_modality(P, hy: Hyper, m, x, mask=None):
    """ModalityEncoder.forward models.py:55-80 for enc in {rnn, none}.  x: [B, hist, d]; mask:
    optional [B, hist] dropout mask already scaled by 1/(1-p) (models.py:56-58)."""
    if mask is not None:
        x = x * mask.unsqueeze(-1)
    e = hy.enc[m]
    if e == "none":
        return x.reshape(x.shape[0], -1)
    if e != "rnn":
        raise NotImplementedError("oracle covers enc in {rnn, none} (SURVEY.md §2 row 5)")
    pre = "feature_encoder.%s_encoder.encoder." % m
    w_ih, w_hh = P[pre + "weight_ih_l0"], P[pre + "weight_hh_l0"]
    b_ih, b_hh = P[pre + "bias_ih_l0"], P[pre + "bias_hh_l0"]
    h = x.new_zeros(x.size(0), hy.enc_hidden[m])
    for s in range(x.size(1)):
        h = torch.gru_cell(x[:, s], h, w_ih, w_hh, b_ih, b_hh)
    return torch.cat([h, h], dim=1)  # seq[:, -1] and h_n[0] are the same tensor (models.py:64)


def conditioning(P, hy: Hyper, data, t, faces, masks=None, ti=None):
    """SeqGlow.create_conditioning (models.py:598-615) + FeatureEncoder.forward (127-145).
    `masks[m]` is [T', B, hist]; `ti` the frame index into it."""
    parts = []
    for m in MODALITIES:
        hist = hy.hist[m]
        if m == "p1_face":
            win = faces[:, t - hist : t]
        else:
            if not hist:
                continue
            win = data[m][:, t - hist + 1 : t + 1]
        mk = masks[m][ti] if (masks is not None and masks.get(m) is not None) else None
        parts.append(encode_modality(P, hy, m, win, mk))
    return torch.cat(parts, dim=1)


# ----------------------------------------------------------------------------- sequence drivers


def seq_forward(P, hy: Hyper, batch, masks=None, scales=None):
    """SeqGlow.forward models.py:534-561.  Returns (z [T',B,C], nll [T',B], loss scalar)."""
    state: dict = {}
    T = batch["p1_face"].shape[1]
    zs, nlls = [], []
    for ti, t in enumerate(range(hy.start_ts, T)):
        cond = conditioning(P, hy, batch, t, batch["p1_face"], masks, ti)
        z, logdet = flow_encode(P, hy, batch["p1_face"][:, t, :], cond, state, scales)
        nlls.append(nll_bits(logdet, z))
        zs.append(z)
    nll = torch.stack(nlls)
    loss = nll.mean(dim=1).sum() / len(zs)
    return torch.stack(zs), nll, loss


def ddi_init(P, hy: Hyper, batch, masks=None):
    """What the first training-mode forward does to ActNorm (modules.py:69-70 reached through
    models.py:546-552 at t=start_ts): layer-by-layer data-dependent init on the first frame."""
    with torch.no_grad():
        state: dict = {}
        t = hy.start_ts
        cond = conditioning(P, hy, batch, t, batch["p1_face"], masks, 0)
        x = batch["p1_face"][:, t, :]
        logdet = torch.zeros_like(x[:, 0])
        for k in range(hy.K):
            actnorm_ddi(P, "glow.flow.layers.%d.actnorm." % k, x, hy.actnorm_scale)
            x, logdet = flow_step(P, hy, k, x, cond, logdet, False, state)


def seq_inference(P, hy: Hyper, data, seq_len, eps=1.0, noise=None, masks=None):
    """SeqGlow.inference models.py:567-596.  `noise` [T',B,C] replaces GaussianDiag.sample
    (already scaled by eps).  Returns [B, seq_len-start_ts, C]."""
    with torch.no_grad():
        state: dict = {}
        faces = data["p1_face"]
        outs = []
        for ti, t in enumerate(range(hy.start_ts, seq_len)):
            cond = conditioning(P, hy, data, t, faces, masks, ti)
            if noise is not None:
                z = noise[ti]
            else:
                z = torch.normal(torch.zeros_like(faces[:, 0, :]), torch.ones_like(faces[:, 0, :]) * eps)
            x, _ = flow_decode(P, hy, z, cond, state)
            faces = torch.cat([faces, x.unsqueeze(1)], dim=1)
            outs.append(x)
        return torch.stack(outs, dim=1)


def seq_invert(P, hy: Hyper, z_seq, data, masks=None):
    """SeqGlow.invert models.py:617-645 (teacher-forced conditioning, given z)."""
    with torch.no_grad():
        state: dict = {}
        rec, loss = [], 0.0
        for ti, z in enumerate(z_seq):
            cond = conditioning(P, hy, data, hy.start_ts + ti, data["p1_face"], masks, ti)
            x, logdet = flow_decode(P, hy, z, cond, state)
            loss = loss + nll_bits(logdet, z).mean()
            rec.append(x)
        return torch.stack(rec), loss / len(rec)


# ----------------------------------------------------------------------------- helpers


def clone_params(sd, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().float()
        if requires_grad and not (k.endswith(".p") or k.endswith(".sign_s")):
            t.requires_grad_(True)
        out[k] = t
    return out


def synthetic_batch(hy: Hyper, B, T, seed=1):
    """N(0,1) from a seeded generator, order p1_face, p2_face, p1_speech, p2_speech."""
    g = torch.Generator().manual_seed(seed)
    return {m: torch.randn(B, T, hy.in_dim[m], generator=g) for m in MODALITIES}


def make_masks(hy: Hyper, B, Tp, seed=3):
    """Dropout masks as `nn.Dropout(p)(ones[B,hist])` would draw them, but from an explicit
    generator so CPU oracle and CUDA path can share them (models.py:56-58)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for m in MODALITIES:
        p = hy.dropout[m]
        if p > 0 and hy.hist[m]:
            keep = (torch.rand(Tp, B, hy.hist[m], generator=g) >= p).float()
            out[m] = keep / (1.0 - p)
        else:
            out[m] = None
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Output side of sampling and training-step glue; test infrastructure only.
def expand_face_dim(seq, exp_dim, jaw_dim, neck_dim):
    """generate_motion_from_model.py:39-51: [B, T, C] -> [B, T, 106]; expression at 0, jaw at 100, neck at 103."""
    out = torch.zeros((seq.size(0), seq.size(1), 106))
    out[:, :, :exp_dim] = seq[:, :, :exp_dim]
    out[:, :, 100:100 + jaw_dim] = seq[:, :, exp_dim:exp_dim + jaw_dim]
    out[:, :, 103:103 + neck_dim] = seq[:, :, exp_dim + jaw_dim:exp_dim + jaw_dim + neck_dim]
    return out


def destandardize_expand(seq, means, stds, exp_dim, jaw_dim, neck_dim):
    """generate_motion_from_model.py:68 then :39-51."""
    return expand_face_dim(seq * stds + means, exp_dim, jaw_dim, neck_dim)


def calc_jerk(x):
    """glow/utils.py:53-58: mean absolute third difference along time of [B, T, C]."""
    d = x[:, 1:] - x[:, :-1]
    a = d[:, 1:] - d[:, :-1]
    j = a[:, 1:] - a[:, :-1]
    return j.abs().mean()


def derange_batch(batch, modalities, permutation):
    """glow/utils.py:85-100 with the permutation drawn by the caller (shuffle_time=False, the only use in the reference)."""
    out = {}
    for m in ("p1_face", "p2_face", "p1_speech", "p2_speech"):
        if m in modalities:
            out[m] = batch[m][permutation]
        elif batch.get(m) is not None:
            out[m] = batch[m]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Input side; test infrastructure only.  The reference's MimicryDataset needs h5py and
# pytorch_lightning, which are absent here, so this restatement is pinned by calling the SAME torch / random primitives the
# reference calls (torch.arange(L).unfold(0, seq_len, 1), random.sample) rather than by running the class: parity for this
# row is anchored on those call sites (mimicry_data_module.py:35-42, 44-78).
def dataset_window_table(segments, seq_len, rng):
    """MimicryDataset.__init__ (mimicry_data_module.py:33-42): [(segment key, [frame indices])], shuffled by random.sample."""
    tmp = []
    for key, seg in enumerate(segments):
        L = len(seg["p1_face"])
        if L >= seq_len:
            for seq in torch.arange(L).unfold(0, seq_len, 1):
                tmp.append((key, seq.int().tolist()))
    return rng.sample(tmp, len(tmp))


def dataset_batch(segments, table, index):
    """__getitem__ (mimicry_data_module.py:44-78) for every item of `index`, stacked as the DataLoader's default collate does."""
    out = {}
    for m in MODALITIES:
        if m in segments[0]:
            out[m] = torch.stack([segments[table[i][0]][m][table[i][1]] for i in index])
    return out
