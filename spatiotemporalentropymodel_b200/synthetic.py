"""Seeded synthetic checkpoints and frames.

The reference ships no checkpoints and the zoo weights need network access, and a default random init is
degenerate for this path (latents ~0.06 std => every symbol 0; ~90 % of the scales under the 0.11 bound). These
builders produce *calibrated* state_dicts with exactly the reference's key set / shapes / dtypes
(`SpatioTemporalPriorModel*.state_dict()` and `models["mbt2018"](quality=4).state_dict()`) from a
seed, so that the reference classes (tests/golden/make_golden.py), the oracle and the CUDA path all load the same
weights without any weight file being stored.  CPU `torch.Generator` streams are platform independent.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

_PEDESTAL = (2.0 ** -18) ** 2


def _u(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _conv(sd, g, name, cout, cin, k, gain=1.0, transposed=False, bias_bound=None):
    fan_in = cin * k * k
    bound = gain * math.sqrt(3.0 / fan_in)  # variance gain^2 / fan_in
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    sd[f"{name}.weight"] = _u(g, shape, bound)
    sd[f"{name}.bias"] = _u(g, (cout,), 0.1 if bias_bound is None else bias_bound)


def _gdn(sd, g, name, c):
    beta = 1.0 + 0.5 * torch.rand(c, generator=g)
    gamma = 0.1 * torch.eye(c) + 0.02 * torch.rand((c, c), generator=g)
    ped = torch.tensor([_PEDESTAL], dtype=torch.float32)
    sd[f"{name}.beta"] = torch.sqrt(torch.max(beta + ped, ped))
    sd[f"{name}.gamma"] = torch.sqrt(torch.max(gamma + ped, ped))
    sd[f"{name}.beta_reparam.pedestal"] = ped.clone()
    sd[f"{name}.beta_reparam.lower_bound.bound"] = torch.tensor([(1e-6 + _PEDESTAL) ** 0.5], dtype=torch.float32)
    sd[f"{name}.gamma_reparam.pedestal"] = ped.clone()
    sd[f"{name}.gamma_reparam.lower_bound.bound"] = torch.tensor([_PEDESTAL ** 0.5], dtype=torch.float32)


def _entropy_bottleneck(sd, g, name, c, init_scale=10.0, filters=(3, 3, 3, 3)):
    f = (1,) + tuple(filters) + (1,)
    scale = init_scale ** (1 / (len(filters) + 1))
    for i in range(len(filters) + 1):
        init = float(np.log(np.expm1(1 / scale / f[i + 1])))
        sd[f"{name}._matrix{i}"] = torch.full((c, f[i + 1], f[i]), init) + 0.05 * torch.randn((c, f[i + 1], f[i]),
                                                                                             generator=g)
        sd[f"{name}._bias{i}"] = _u(g, (c, f[i + 1], 1), 0.5)
        if i < len(filters):
            sd[f"{name}._factor{i}"] = 0.1 * torch.randn((c, f[i + 1], 1), generator=g)
    q = torch.tensor([-init_scale, 0.0, init_scale]).repeat(c, 1, 1)
    q[:, 0, 1] = 0.3 * torch.randn(c, generator=g)
    sd[f"{name}.quantiles"] = q
    sd[f"{name}._offset"] = torch.IntTensor()
    sd[f"{name}._quantized_cdf"] = torch.IntTensor()
    sd[f"{name}._cdf_length"] = torch.IntTensor()
    target = float(np.log(2 / 1e-9 - 1))
    sd[f"{name}.target"] = torch.tensor([-target, 0.0, target])
    sd[f"{name}.likelihood_lower_bound.bound"] = torch.tensor([1e-9])


def _gaussian_conditional(sd, name="gaussian_conditional"):
    sd[f"{name}._offset"] = torch.IntTensor()
    sd[f"{name}._quantized_cdf"] = torch.IntTensor()
    sd[f"{name}._cdf_length"] = torch.IntTensor()
    sd[f"{name}.scale_table"] = torch.Tensor()
    sd[f"{name}.scale_bound"] = torch.tensor([0.11])
    sd[f"{name}.likelihood_lower_bound.bound"] = torch.tensor([1e-9])
    sd[f"{name}.lower_bound_scale.bound"] = torch.tensor([0.11])


def _mask_a(cout, cin, k=5):
    m = torch.ones((cout, cin, k, k))
    m[:, :, k // 2, k // 2:] = 0
    m[:, :, k // 2 + 1:] = 0
    return m


def make_iframe_state_dict(seed: int = 0, N: int = 192, M: int = 192) -> Dict[str, Tensor]:
    """Key set of JointAutoregressiveHierarchicalPriors(N, M) (priors.py:418-475), i.e. models["mbt2018"](4)."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd: Dict[str, Tensor] = {}
    _entropy_bottleneck(sd, g, "entropy_bottleneck", N)
    # analysis: gains chosen so activations stay O(1) through GDN and the latent has std ~ 3
    _conv(sd, g, "g_a.0", N, 3, 5, gain=3.0)
    _gdn(sd, g, "g_a.1", N)
    _conv(sd, g, "g_a.2", N, N, 5, gain=2.0)
    _gdn(sd, g, "g_a.3", N)
    _conv(sd, g, "g_a.4", N, N, 5, gain=2.0)
    _gdn(sd, g, "g_a.5", N)
    _conv(sd, g, "g_a.6", M, N, 5, gain=6.0)
    # synthesis: keep x_hat inside (0, 1) without saturating everywhere
    _conv(sd, g, "g_s.0", N, M, 5, gain=0.6, transposed=True)
    _gdn(sd, g, "g_s.1", N)
    _conv(sd, g, "g_s.2", N, N, 5, gain=1.0, transposed=True)
    _gdn(sd, g, "g_s.3", N)
    _conv(sd, g, "g_s.4", N, N, 5, gain=1.0, transposed=True)
    _gdn(sd, g, "g_s.5", N)
    _conv(sd, g, "g_s.6", 3, N, 5, gain=0.5, transposed=True)
    sd["g_s.6.bias"] = torch.tensor([0.45, 0.5, 0.55])
    _conv(sd, g, "h_a.0", N, M, 3)
    _conv(sd, g, "h_a.2", N, N, 5)
    _conv(sd, g, "h_a.4", N, N, 5)
    _conv(sd, g, "h_s.0", M, N, 5, transposed=True)
    _conv(sd, g, "h_s.2", M * 3 // 2, M, 5, transposed=True)
    _conv(sd, g, "h_s.4", M * 2, M * 3 // 2, 3)
    _gaussian_conditional(sd)
    _conv(sd, g, "entropy_parameters.0", M * 10 // 3, M * 12 // 3, 1)
    _conv(sd, g, "entropy_parameters.2", M * 8 // 3, M * 10 // 3, 1)
    _conv(sd, g, "entropy_parameters.4", M * 6 // 3, M * 8 // 3, 1)
    _conv(sd, g, "context_prediction", 2 * M, M, 5)
    sd["context_prediction.mask"] = _mask_a(2 * M, M)
    return sd


STEM_VARIANTS = (
    "SpatioTemporalPriorModel",
    "SpatioTemporalPriorModel_Res",
    "SpatioTemporalPriorModelWithoutSPM",
    "SpatioTemporalPriorModelWithoutTPM",
    "SpatioTemporalPriorModelWithoutSPMTPM",
)


def variant_flags(variant: str) -> Tuple[bool, bool, bool]:
    """(has_tpm, has_spm, residual)"""
    if variant not in STEM_VARIANTS:
        raise ValueError(f"unknown STEM variant {variant!r}")
    has_tpm = variant not in ("SpatioTemporalPriorModelWithoutTPM", "SpatioTemporalPriorModelWithoutSPMTPM")
    has_spm = variant not in ("SpatioTemporalPriorModelWithoutSPM", "SpatioTemporalPriorModelWithoutSPMTPM")
    return has_tpm, has_spm, variant == "SpatioTemporalPriorModel_Res"


def make_stem_state_dict(variant: str = "SpatioTemporalPriorModel", seed: int = 0, in_channels: int = 192,
                         eb_channels: int = 256) -> Dict[str, Tensor]:
    """Key set of the five classes in compressai/models/spatiotemporalpriors.py (ctor :516)."""
    has_tpm, has_spm, _ = variant_flags(variant)
    g = torch.Generator().manual_seed(2000 + seed)
    C = in_channels
    sd: Dict[str, Tensor] = {}
    _entropy_bottleneck(sd, g, "entropy_bottleneck", eb_channels)
    if has_tpm:
        _conv(sd, g, "TPM.0", 256, C, 5, gain=1.4)
        _conv(sd, g, "TPM.2", 320, 256, 5, gain=1.4)
        _conv(sd, g, "TPM.4", 2 * C, 320, 5, gain=1.4)
    _conv(sd, g, "HE.0", 256, 2 * C, 3, gain=1.4)
    _conv(sd, g, "HE.2", 256, 256, 5, gain=1.4)
    _conv(sd, g, "HE.4", eb_channels if has_tpm or has_spm else 256, 256, 5, gain=2.0)
    _conv(sd, g, "HD.0", 256, eb_channels if has_tpm or has_spm else 256, 5, gain=1.4, transposed=True)
    _conv(sd, g, "HD.2", 256, 256, 5, gain=1.4, transposed=True)
    _conv(sd, g, "HD.4", 2 * C, 256, 3, gain=1.4)
    if has_spm:
        _conv(sd, g, "context_prediction", 2 * C, C, 5, gain=1.0)
        sd["context_prediction.mask"] = _mask_a(2 * C, C)
    n_prior = (1 + int(has_tpm) + int(has_spm)) * 2 * C
    _conv(sd, g, "EPM.0", 768, n_prior, 1, gain=1.4)
    _conv(sd, g, "EPM.2", 576, 768, 1, gain=1.4)
    _conv(sd, g, "EPM.4", 2 * C, 576, 1, gain=0.7)
    # scales: log-spaced biases so sigma covers the whole scale table including the 0.11 clamp;
    # means: small so that |y - mu| stays a few sigma for most elements
    b = sd["EPM.4.bias"]
    b[:C] = torch.exp(torch.linspace(math.log(0.05), math.log(64.0), C))
    sd["EPM.4.weight"][:C] *= 0.3
    _gaussian_conditional(sd)
    return sd


def make_frames(n_frames: int, height: int, width: int, seed: int = 1234) -> Tensor:
    """Smooth, temporally correlated frames in [0, 1]: bicubic-upsampled noise, shifted per
    frame, plus 0.02 sigma noise. Returns (n_frames, 3, H, W) fp32."""
    g = torch.Generator().manual_seed(seed)
    hb, wb = max(height // 16, 2), max(width // 16, 2)
    base = F.interpolate(torch.rand((1, 3, hb, wb), generator=g), size=(height, width), mode="bicubic",
                         align_corners=False).clamp(0, 1)
    frames = []
    for t in range(n_frames):
        f = torch.roll(base, shifts=(t, 2 * t), dims=(2, 3))
        f = (f + 0.02 * torch.randn(f.shape, generator=g)).clamp(0, 1)
        frames.append(f)
    return torch.cat(frames, 0)


def make_latent(n: int, c: int, h: int, w: int, seed: int = 4321, std: float = 3.0) -> Tensor:
    """Integer-valued stand-in for the previous decoded latent y_hat (the I-frame codec's output)."""
    g = torch.Generator().manual_seed(seed)
    return torch.round(std * torch.randn((n, c, h, w), generator=g))
