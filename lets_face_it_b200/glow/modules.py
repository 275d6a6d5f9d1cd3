"""Flow primitives with the reference's module API, backed by the sm_100a kernels.

Same class names, constructor signatures, parameter/buffer names and shapes as
`code/glow_pytorch/glow/modules.py` of the reference (so its checkpoints load), but
`forward` calls liblfi_b200.so through `_cabi` — CUDA tensors only, no torch fallback.
Gradients flow through the fused `SeqGlow.forward` path and through `FlowStep / FlowNet / Glow.forward` (one autograd node
per flow step and frame, models.py: `_FlowStepFn`); the stand-alone `ActNorm2d / InvertibleConv1x1 / LinearZeros` calls below
(the reference's test_modules.py:9-27 round trips) are inference-style.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import torch
import torch.nn as nn

from .. import _cabi as cabi


def _f32c(t, device=None):
    t = t.detach()
    if device is not None:
        t = t.to(device)
    return t.to(torch.float32).contiguous()


class ActNorm2d(nn.Module):
    """y = (x + bias) * exp(logs), per channel on [B, C]; log-det term C * sum(logs) (the reference
    multiplies by input.size(1), modules.py:62).  Data-dependent init on the first training call
    (modules.py:32-43)."""

    def __init__(self, num_features, scale=1.0):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(1, num_features))
        self.logs = nn.Parameter(torch.zeros(1, num_features))
        self.num_features = num_features
        self.scale = float(scale)
        self.inited = False

    def initialize_parameters(self, input):
        if not self.training:
            return
        assert input.device == self.bias.device
        with torch.no_grad():
            x = input.detach().float()
            bias = -x.mean(dim=0, keepdim=True)
            var = ((x + bias) ** 2).mean(dim=0, keepdim=True)
            logs = torch.log(self.scale / (torch.sqrt(var) + 1e-6))
            self.bias.data.copy_(bias)
            self.logs.data.copy_(logs)
            self.inited = True

    def forward(self, input, logdet=None, reverse=False):
        if not self.inited:
            self.initialize_parameters(input)
        x = _f32c(input)
        y = torch.empty_like(x)
        B, C = x.shape
        cabi.check(cabi.lib().lfi_actnorm(cabi.ptr(x), cabi.ptr(_f32c(self.bias, x.device)), cabi.ptr(_f32c(self.logs, x.device)),
                                          cabi.ptr(y), B, C, int(bool(reverse)), cabi.stream_ptr()), "lfi_actnorm")
        if logdet is not None:
            dlogdet = self.logs.detach().sum().to(x.device) * C
            logdet = logdet - dlogdet if reverse else logdet + dlogdet
        return y, logdet


class LinearZeros(nn.Linear):
    """(h W^T + b) * exp(logs * logscale_factor), zero initialised (modules.py:83-95)."""

    def __init__(self, in_channels, out_channels, logscale_factor=3):
        super().__init__(in_channels, out_channels)
        self.logscale_factor = logscale_factor
        self.logs = nn.Parameter(torch.zeros(out_channels))
        self.weight.data.zero_()
        self.bias.data.zero_()

    def forward(self, input):
        x = _f32c(input)
        B = x.shape[0]
        out = torch.empty(B, self.out_features, dtype=torch.float32, device=x.device)
        cabi.check(cabi.lib().lfi_matmul(cabi.ptr(x), cabi.ptr(_f32c(self.weight, x.device)), cabi.ptr(_f32c(self.bias, x.device)),
                                         cabi.ptr(out), B, self.out_features, self.in_features, 1, cabi.stream_ptr()), "lfi_matmul")
        return out * torch.exp(self.logs.detach().to(x.device) * self.logscale_factor)


class InvertibleConv1x1(nn.Module):
    """z = x @ W with W = P (L*mask + I)(U*mask^T + diag(sign_s e^{log_s})) when LU_decomposed
    (modules.py:122-194); log-det term C * sum(log_s).  Init: QR of a numpy Gaussian + scipy LU,
    drawn exactly as the reference draws it (modules.py:126-143)."""

    def __init__(self, num_channels, LU_decomposed=False):
        super().__init__()
        w_shape = [num_channels, num_channels]
        w_init = np.linalg.qr(np.random.randn(*w_shape))[0].astype(np.float32)
        if not LU_decomposed:
            self.weight = nn.Parameter(torch.Tensor(w_init))
        else:
            np_p, np_l, np_u = scipy.linalg.lu(w_init)
            np_s = np.diag(np_u)
            self.register_buffer("p", torch.Tensor(np_p.astype(np.float32)))
            self.register_buffer("sign_s", torch.Tensor(np.sign(np_s).astype(np.float32)))
            self.l = nn.Parameter(torch.Tensor(np_l.astype(np.float32)))
            self.log_s = nn.Parameter(torch.Tensor(np.log(np.abs(np_s)).astype(np.float32)))
            self.u = nn.Parameter(torch.Tensor(np.triu(np_u, k=1).astype(np.float32)))
        self.w_shape = w_shape
        self.LU = LU_decomposed

    def get_weight(self, input, reverse):
        C = self.w_shape[0]
        dev = input.device
        if not self.LU:
            w = _f32c(self.weight, dev)
            dlogdet = torch.slogdet(w)[1] * input.size(1)
            if reverse:
                w = torch.inverse(w.double()).float()
            return w, dlogdet
        L = cabi.lib()
        w = torch.empty(C, C, dtype=torch.float32, device=dev)
        winv = torch.empty(C, C, dtype=torch.float32, device=dev) if reverse else None
        ws = torch.empty(L.lfi_invconv_ws_bytes(1, C), dtype=torch.uint8, device=dev)
        cabi.check(L.lfi_invconv_compose(1, C, cabi.ptr(_f32c(self.p, dev)), cabi.ptr(_f32c(self.l, dev)), cabi.ptr(_f32c(self.u, dev)),
                                         cabi.ptr(_f32c(self.log_s, dev)), cabi.ptr(_f32c(self.sign_s, dev)), cabi.ptr(w),
                                         cabi.ptr(winv), ws.data_ptr(), ws.numel(), cabi.stream_ptr()), "lfi_invconv_compose")
        dlogdet = self.log_s.detach().sum().to(dev) * input.size(1)
        return (winv if reverse else w), dlogdet

    def forward(self, input, logdet=None, reverse=False):
        x = _f32c(input)
        weight, dlogdet = self.get_weight(x, reverse)
        B, C = x.shape
        z = torch.empty_like(x)
        cabi.check(cabi.lib().lfi_matmul(cabi.ptr(x), cabi.ptr(weight.contiguous()), None, cabi.ptr(z), B, C, C, 0, cabi.stream_ptr()),
                   "lfi_matmul")
        if logdet is not None:
            logdet = logdet - dlogdet if reverse else logdet + dlogdet
        return z, logdet


class GaussianDiag:
    """Standard-normal prior helpers (modules.py:197-235)."""

    Log2PI = float(np.log(2 * np.pi))

    @staticmethod
    def likelihood_simplified(x):
        return -0.5 * ((x ** 2) + GaussianDiag.Log2PI)

    @staticmethod
    def logp_simplified(x):
        return torch.sum(GaussianDiag.likelihood_simplified(x), dim=1)

    @staticmethod
    def likelihood(mean, logs, x):
        return -0.5 * (logs * 2.0 + ((x - mean) ** 2) / torch.exp(logs * 2.0) + GaussianDiag.Log2PI)

    @staticmethod
    def logp(mean, logs, x):
        return torch.sum(GaussianDiag.likelihood(mean, logs, x), dim=1)

    @staticmethod
    def sample(output_shape, eps_std=1):
        return torch.normal(mean=torch.zeros_like(output_shape), std=torch.ones_like(output_shape) * eps_std)

    @staticmethod
    def nll_bits(z, logdet):
        """-(logdet + logp_simplified(z)) / ln 2 on the device (SeqGlow.loss, models.py:563-565)."""
        z = _f32c(z)
        ld = _f32c(logdet)
        out = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        cabi.check(cabi.lib().lfi_nll(cabi.ptr(z), cabi.ptr(ld), cabi.ptr(out), z.shape[0], z.shape[1], cabi.stream_ptr()), "lfi_nll")
        return out
