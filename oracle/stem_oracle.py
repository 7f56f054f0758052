"""CPU oracle for the STEM P-frame hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product package ``spatiotemporalentropymodel_b200`` never does (it fails loudly when
its CUDA library is missing instead of falling back to anything here).

What it is: a functional restatement (plain ``torch`` CPU ops on a ``state_dict``; no ``nn.Module``) of the
arithmetic of mmSir/SpatioTemporalEntropyModel's P-frame path.  The reference's own arithmetic is PyTorch ATen
(``requirements.txt:93`` pins torch==1.7.1; this image has 2.11), so "the reference's CPU algorithm" *is*
``F.conv2d`` & co. on fp32 tensors; every function cites the reference lines it follows.

Parity pinning: the reference ships no golden vectors for this path.  The oracle is pinned by
``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py`` which imports the *reference classes
themselves* (from a throw-away built copy of /root/reference), loads the seeded synthetic checkpoints of
``spatiotemporalentropymodel_b200.synthetic`` into them and records their outputs; ``tests/test_oracle.py``
checks this module against those fixtures and against the reference's own known-answer tests
(``compressai_tests/test_layers.py:118-143`` GDN/IGDN closed form, ``test_entropy_models.py:58-71,249-286``
rounding semantics, ``test_models.py:173-181`` scale table ends).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64  # spatiotemporalpriors.py:22-25
LIKELIHOOD_BOUND = 1e-9                                # entropy_models.py:79
SCALE_BOUND = 0.11                                     # entropy_models.py:484

VARIANTS = (
    "SpatioTemporalPriorModel",
    "SpatioTemporalPriorModel_Res",
    "SpatioTemporalPriorModelWithoutSPM",
    "SpatioTemporalPriorModelWithoutTPM",
    "SpatioTemporalPriorModelWithoutSPMTPM",
)


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS) -> Tensor:
    """spatiotemporalpriors.py:27-30"""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


# ----------------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------------
def lower_bound(x: Tensor, bound: float) -> Tensor:
    """ops/bound_ops.py:50-53 forward: torch.max(x, bound)"""
    return torch.max(x, torch.tensor([bound], dtype=x.dtype))


def nonneg_reparam(p: Tensor, minimum: float = 0.0, offset: float = 2 ** -18) -> Tensor:
    """ops/parametrizers.py:27-45: max(p, sqrt(minimum + offset^2))^2 - offset^2 (all in fp32 tensors)."""
    pedestal = torch.tensor([offset ** 2], dtype=torch.float32)
    bound = (minimum + offset ** 2) ** 0.5
    out = lower_bound(p, bound)
    return out ** 2 - pedestal


def gdn(x: Tensor, beta_p: Tensor, gamma_p: Tensor, inverse: bool) -> Tensor:
    """layers/gdn.py:52-67 (beta_min = 1e-6, gdn.py:34)."""
    c = x.shape[1]
    beta = nonneg_reparam(beta_p, minimum=1e-6)
    gamma = nonneg_reparam(gamma_p).reshape(c, c, 1, 1)
    norm = F.conv2d(x ** 2, gamma, beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


def g_a(x: Tensor, sd: SD, prefix: str = "g_a") -> Tensor:
    """priors.py:421-429 with conv() = Conv2d(k5, s2, p2) (models/utils.py:112-119)."""
    h = x
    for i in (0, 2, 4, 6):
        h = F.conv2d(h, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"], stride=2, padding=2)
        if i < 6:
            h = gdn(h, sd[f"{prefix}.{i + 1}.beta"], sd[f"{prefix}.{i + 1}.gamma"], inverse=False)
    return h


def g_s(y_hat: Tensor, sd: SD, prefix: str = "g_s", clamp: bool = True) -> Tensor:
    """priors.py:431-439 with deconv() = ConvTranspose2d(k5, s2, p2, op1) (models/utils.py:122-130);
    getX clamps to [0, 1] (priors.py:397-402)."""
    h = y_hat
    for i in (0, 2, 4, 6):
        h = F.conv_transpose2d(h, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"], stride=2, padding=2,
                               output_padding=1)
        if i < 6:
            h = gdn(h, sd[f"{prefix}.{i + 1}.beta"], sd[f"{prefix}.{i + 1}.gamma"], inverse=True)
    return h.clamp(0, 1) if clamp else h


def lrelu(x: Tensor) -> Tensor:
    return F.leaky_relu(x, 0.01)  # nn.LeakyReLU() default slope


# ----------------------------------------------------------------------------------------------------
# entropy models
# ----------------------------------------------------------------------------------------------------
def quantize_dequantize(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """entropy_models.py:136-146 ("dequantize" mode): round(x - m) + m, torch.round = half to even."""
    out = x.clone()
    if means is not None:
        out -= means
    out = torch.round(out)
    if means is not None:
        out += means
    return out


def quantize_symbols(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """entropy_models.py:136-150 ("symbols" mode)."""
    out = x.clone()
    if means is not None:
        out -= means
    return torch.round(out).int()


def _std_cumulative(v: Tensor) -> Tensor:
    """entropy_models.py:521-526"""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)


def gaussian_likelihood(y_hat: Tensor, scales: Tensor, means: Optional[Tensor]) -> Tensor:
    """entropy_models.py:570-586 (without the final likelihood bound)."""
    values = y_hat - means if means is not None else y_hat
    scales = lower_bound(scales, SCALE_BOUND)
    values = torch.abs(values)
    upper = _std_cumulative((0.5 - values) / scales)
    lower = _std_cumulative((-0.5 - values) / scales)
    return upper - lower


def gaussian_conditional_forward(y: Tensor, scales: Tensor, means: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """entropy_models.py:588-596, eval mode."""
    y_hat = quantize_dequantize(y, means)
    lik = gaussian_likelihood(y_hat, scales, means)
    return y_hat, lower_bound(lik, LIKELIHOOD_BOUND)


def build_indexes(scales: Tensor, table: Optional[Tensor] = None) -> Tensor:
    """entropy_models.py:598-604"""
    table = get_scale_table() if table is None else table
    scales = lower_bound(scales, SCALE_BOUND)
    idx = scales.new_full(scales.size(), len(table) - 1).int()
    for s in table[:-1]:
        idx -= (scales <= s).int()
    return idx


def eb_logits_cumulative(v: Tensor, sd: SD, prefix: str) -> Tensor:
    """entropy_models.py:388-407; v: (C, 1, N)"""
    logits = v
    for i in range(5):
        logits = torch.matmul(F.softplus(sd[f"{prefix}._matrix{i}"]), logits)
        logits = logits + sd[f"{prefix}._bias{i}"]
        if i < 4:
            logits = logits + torch.tanh(sd[f"{prefix}._factor{i}"]) * torch.tanh(logits)
    return logits


def entropy_bottleneck_forward(z: Tensor, sd: SD, prefix: str = "entropy_bottleneck") -> Tuple[Tensor, Tensor]:
    """entropy_models.py:424-452 eval mode (+ :410-422 likelihood, medians :337-339)."""
    x = z.permute(1, 2, 3, 0).contiguous()
    shape = x.size()
    values = x.reshape(x.size(0), 1, -1)
    medians = sd[f"{prefix}.quantiles"][:, :, 1:2]
    outputs = quantize_dequantize(values, medians)
    lower = eb_logits_cumulative(outputs - 0.5, sd, prefix)
    upper = eb_logits_cumulative(outputs + 0.5, sd, prefix)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = lower_bound(lik, LIKELIHOOD_BOUND)
    outputs = outputs.reshape(shape).permute(3, 0, 1, 2).contiguous()
    lik = lik.reshape(shape).permute(3, 0, 1, 2).contiguous()
    return outputs, lik


# ----------------------------------------------------------------------------------------------------
# STEM sub-networks (spatiotemporalpriors.py:523-554)
# ----------------------------------------------------------------------------------------------------
def _seq_conv(x: Tensor, sd: SD, name: str, specs) -> Tensor:
    """specs: list of (index, kind, stride, padding); LeakyReLU between layers."""
    for n, (i, kind, stride, pad) in enumerate(specs):
        w, b = sd[f"{name}.{i}.weight"], sd[f"{name}.{i}.bias"]
        if kind == "conv":
            x = F.conv2d(x, w, b, stride=stride, padding=pad)
        else:
            x = F.conv_transpose2d(x, w, b, stride=stride, padding=pad, output_padding=1)
        if n + 1 < len(specs):
            x = lrelu(x)
    return x


def TPM(y_cond: Tensor, sd: SD) -> Tensor:
    return _seq_conv(y_cond, sd, "TPM", [(0, "conv", 1, 2), (2, "conv", 1, 2), (4, "conv", 1, 2)])


def HE(y_cat: Tensor, sd: SD) -> Tensor:
    return _seq_conv(y_cat, sd, "HE", [(0, "conv", 1, 1), (2, "conv", 2, 2), (4, "conv", 2, 2)])


def HD(z_hat: Tensor, sd: SD) -> Tensor:
    return _seq_conv(z_hat, sd, "HD", [(0, "deconv", 2, 2), (2, "deconv", 2, 2), (4, "conv", 1, 1)])


def EPM(x: Tensor, sd: SD) -> Tensor:
    return _seq_conv(x, sd, "EPM", [(0, "conv", 1, 0), (2, "conv", 1, 0), (4, "conv", 1, 0)])


def masked_weight(sd: SD) -> Tensor:
    """layers/layers.py:38-46: weight *= mask, mask 'A' zeroes (h//2, w//2:) and rows below."""
    w = sd["context_prediction.weight"]
    mask = torch.ones_like(w)
    _, _, kh, kw = w.shape
    mask[:, :, kh // 2, kw // 2:] = 0
    mask[:, :, kh // 2 + 1:] = 0
    return w * mask


def context_prediction(y_hat: Tensor, sd: SD) -> Tensor:
    return F.conv2d(y_hat, masked_weight(sd), sd["context_prediction.bias"], padding=2)


def stem_forward(variant: str, y_cur: Tensor, y_cond: Tensor, sd: SD, return_params: bool = False):
    """forward() of the five classes, eval mode:
    full :561-585, _Res :845-868, WithoutSPM :176-194, WithoutTPM :291-309, WithoutSPMTPM :70-83."""
    assert variant in VARIANTS, variant
    z = HE(torch.cat([y_cur, y_cond], 1), sd)
    z_hat, z_lik = entropy_bottleneck_forward(z, sd)
    hp = HD(z_hat, sd)
    has_tpm = variant not in ("SpatioTemporalPriorModelWithoutTPM", "SpatioTemporalPriorModelWithoutSPMTPM")
    has_spm = variant not in ("SpatioTemporalPriorModelWithoutSPM", "SpatioTemporalPriorModelWithoutSPMTPM")
    res = variant == "SpatioTemporalPriorModel_Res"
    parts = []
    if has_tpm:
        parts.append(TPM(y_cond, sd))
    parts.append(hp)
    target = y_cur - y_cond if res else y_cur
    q = None
    if has_spm:
        q = quantize_dequantize(target)           # round(y) / round(y - y_cond), means=None
        parts.append(context_prediction(q, sd))
    gp = EPM(torch.cat(parts, 1), sd)
    scales, means = gp.chunk(2, 1)
    y_hat_gc, y_lik = gaussian_conditional_forward(target, scales, means)
    if has_spm:
        y_hat = q + y_cond if res else q
    else:
        y_hat = y_hat_gc
    out = {"y_hat": y_hat, "likelihoods": {"y": y_lik, "z": z_lik}}
    if return_params:
        out["scales"], out["means"], out["z_hat"], out["z"] = scales, means, z_hat, z
    return out


# ----------------------------------------------------------------------------------------------------
# whole P-frame (stem/evalSTEM.py:93-154)
# ----------------------------------------------------------------------------------------------------
def pad_to_64(x: Tensor) -> Tuple[Tensor, Tuple[int, int, int, int]]:
    """evalSTEM.py:96-109: centred zero padding to a multiple of 64."""
    h, w = x.size(2), x.size(3)
    p = 64
    new_h, new_w = (h + p - 1) // p * p, (w + p - 1) // p * p
    left = (new_w - w) // 2
    right = new_w - w - left
    top = (new_h - h) // 2
    bottom = new_h - h - top
    return F.pad(x, (left, right, top, bottom), mode="constant", value=0), (left, right, top, bottom)


def pframe_forward(x: Tensor, y_cond: Tensor, sd_i: SD, sd_stem: SD, variant: str,
                   return_params: bool = False) -> Dict[str, Tensor]:
    """evalSTEM.py:93-154 minus the entropy coder: pad -> g_a -> STEM forward -> g_s on the forward pass's
    y_hat -> crop -> estimated bpp (:133-136) and PSNR (:29-31,146)."""
    x_pad, (l, r, t, b) = pad_to_64(x)
    y = g_a(x_pad, sd_i)
    out = stem_forward(variant, y, y_cond, sd_stem, return_params=return_params)
    x_hat = g_s(out["y_hat"], sd_i)
    x_hat = F.pad(x_hat, (-l, -r, -t, -b))
    n, _, h, w = x.shape
    num_pixels = h * w  # per frame
    bits_y = (torch.log(out["likelihoods"]["y"]).flatten(1).sum(1) / -math.log(2))
    bits_z = (torch.log(out["likelihoods"]["z"]).flatten(1).sum(1) / -math.log(2))
    mse = ((x - x_hat) ** 2).flatten(1).mean(1)
    extra = {k: out[k] for k in ("scales", "means", "z_hat") if k in out}
    return {
        **extra, "y": y, "y_hat": out["y_hat"], "x_hat": x_hat,
        "lik_y": out["likelihoods"]["y"], "lik_z": out["likelihoods"]["z"],
        "bpp": (bits_y + bits_z) / num_pixels, "bpp_y": bits_y / num_pixels, "bpp_z": bits_z / num_pixels,
        "psnr": -10 * torch.log10(mse), "mse": mse,
    }


def gop_forward(frames: Tensor, y_cond0: Tensor, sd_i: SD, sd_stem: SD, variant: str, return_params: bool = False):
    """evalSTEM.py:184-209 P-frame loop: y_conditioned <- y_hat of the previous frame."""
    outs, y_cond = [], y_cond0
    for t in range(frames.shape[0]):
        o = pframe_forward(frames[t:t + 1], y_cond, sd_i, sd_stem, variant, return_params=return_params)
        outs.append(o)
        y_cond = o["y_hat"]
    return outs


# ----------------------------------------------------------------------------------------------------
# update(): CDF tables (entropy_models.py:543-568, :170-178; cpp_exts/ops/ops.cpp:24-81)
# ----------------------------------------------------------------------------------------------------
def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """ops.cpp:24-81 restated with numpy integers."""
    pmf = np.asarray(pmf, dtype=np.float32)
    n = len(pmf) + 1
    cdf = np.zeros(n, dtype=np.int64)
    cdf[1:] = np.round(pmf * np.float32(1 << precision)).astype(np.int64)  # std::round on fp32 product
    # np.round is half-to-even, std::round is half-away-from-zero: fix exact .5 cases
    prod = pmf * np.float32(1 << precision)
    frac = prod - np.floor(prod)
    cdf[1:] = np.where(frac == 0.5, np.floor(prod) + 1, np.round(prod)).astype(np.int64)
    total = int(cdf.sum())
    cdf = ((1 << precision) * cdf) // total
    cdf = np.cumsum(cdf)
    cdf[-1] = 1 << precision
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best = None, -1
            for j in range(n - 1):
                f = cdf[j + 1] - cdf[j]
                if f > 1 and (best_freq is None or f < best_freq):
                    best_freq, best = f, j
            assert best != -1
            if best < i:
                cdf[best + 1:i + 1] -= 1
            else:
                cdf[i + 1:best + 1] += 1
    return cdf.astype(np.int32)


def gaussian_conditional_tables(scale_table: Optional[Tensor] = None, tail_mass: float = 1e-9):
    """entropy_models.py:543-568 -> (_quantized_cdf, _offset, _cdf_length)"""
    import scipy.stats
    table = get_scale_table() if scale_table is None else scale_table
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    pmf_center = torch.ceil(table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length).item())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = table.unsqueeze(1).float()
    upper = _std_cumulative((0.5 - samples) / scale)
    lower = _std_cumulative((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i in range(len(pmf_length)):
        prob = torch.cat((pmf[i, : pmf_length[i]], tail[i]), dim=0)
        c = torch.from_numpy(pmf_to_quantized_cdf(prob.numpy(), 16))
        cdf[i, : c.numel()] = c
    return cdf, -pmf_center, pmf_length + 2


# ----------------------------------------------------------------------------------------------------
# autoregressive coding: the per-position scans of the reference, restated
# ----------------------------------------------------------------------------------------------------
def ar_scan(target: Optional[Tensor], priors: Tensor, sd: SD, head: str = "EPM",
            symbols: Optional[Tensor] = None, table: Optional[Tensor] = None):
    """_compress_ar (spatiotemporalpriors.py:633-678, _Res :915-961, priors.py:556-600) when ``target`` is given,
    _decompress_ar (:729-768 / :1016-1055 / priors.py:651-684) when ``symbols`` (1-D, (h, w, c) order) is given.
    target: (1, C, H, W) quantity being coded (y, or y - y_cond); priors: (1, P, H, W) = cat of the EPM inputs that
    precede the context features (tp | hp, hp, or h_s(z_hat)). Raster scan, one latent position at a time:
    5x5 masked-conv crop -> 1x1 head -> (scales, means) -> build_indexes, round(t - mu), t_hat = sym + mu.
    Returns (t_hat (1, C, H, W), symbols (H*W*C,), indexes (H*W*C,)) in the stream order of the reference."""
    assert (target is None) != (symbols is None)
    if table is None:
        table = get_scale_table()
    w_ctx, b_ctx = masked_weight(sd), sd["context_prediction.bias"]
    _, _, H, W = priors.shape
    C = w_ctx.shape[1]
    pad = 2
    t_hat = torch.zeros((1, C, H + 2 * pad, W + 2 * pad))
    if target is not None:
        t_hat = F.pad(target, (pad, pad, pad, pad)).clone()
    sym_out, idx_out = [], []
    k = 0
    for h in range(H):
        for w in range(W):
            crop = t_hat[:, :, h:h + 5, w:w + 5]
            ctx_p = F.conv2d(crop, w_ctx, bias=b_ctx)
            p = priors[:, :, h:h + 1, w:w + 1]
            g = _seq_conv(torch.cat((p, ctx_p), dim=1), sd, head, [(0, "conv", 1, 0), (2, "conv", 1, 0), (4, "conv", 1, 0)])
            g = g.squeeze(3).squeeze(2)
            scales, means = g.chunk(2, 1)
            idx = build_indexes(scales, table)
            if target is not None:
                s = quantize_symbols(crop[:, :, pad, pad], means)
            else:
                s = symbols[k:k + C].reshape(1, C).to(torch.int32)
            k += C
            t_hat[:, :, h + pad, w + pad] = s.to(means.dtype) + means
            sym_out.append(s.reshape(-1))
            idx_out.append(idx.reshape(-1))
    t_hat = t_hat[:, :, pad:-pad, pad:-pad].contiguous()
    return t_hat, torch.cat(sym_out).to(torch.int32), torch.cat(idx_out).to(torch.int32)


def eb_quantize(z: Tensor, sd: SD, prefix: str = "entropy_bottleneck") -> Tensor:
    """EntropyBottleneck.compress -> decompress is lossless: z_hat = round(z - median) + median
    (entropy_models.py:454-471 with the medians of :337-339)."""
    med = sd[f"{prefix}.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
    return torch.round(z - med) + med


def stem_ar_code(variant: str, y_cur: Tensor, y_cond: Tensor, sd: SD):
    """compress() of the SPM variants (spatiotemporalpriors.py:588-631, _Res :871-913, WithoutTPM :311-352):
    -> dict(y_hat = what decompress() returns, symbols, indexes, z_hat)."""
    has_tpm = variant != "SpatioTemporalPriorModelWithoutTPM"
    res = variant == "SpatioTemporalPriorModel_Res"
    z_hat = eb_quantize(HE(torch.cat([y_cur, y_cond], 1), sd), sd)
    parts = ([TPM(y_cond, sd)] if has_tpm else []) + [HD(z_hat, sd)]
    target = y_cur - y_cond if res else y_cur
    t_hat, sym, idx = ar_scan(target, torch.cat(parts, 1), sd)
    return {"y_hat": t_hat + y_cond if res else t_hat, "symbols": sym, "indexes": idx, "z_hat": z_hat}


def iframe_hyper(y: Tensor, sd: SD) -> Tensor:
    return _seq_conv(y, sd, "h_a", [(0, "conv", 1, 1), (2, "conv", 2, 2), (4, "conv", 2, 2)])


def iframe_hyper_synthesis(z_hat: Tensor, sd: SD) -> Tensor:
    return _seq_conv(z_hat, sd, "h_s", [(0, "deconv", 2, 2), (2, "deconv", 2, 2), (4, "conv", 1, 1)])


def iframe_forward(x: Tensor, sd: SD) -> Dict[str, Tensor]:
    """JointAutoregressiveHierarchicalPriors.forward, eval mode (priors.py:477-508)."""
    y = g_a(x, sd)
    z = iframe_hyper(y, sd)
    z_hat, z_lik = entropy_bottleneck_forward(z, sd)
    params = iframe_hyper_synthesis(z_hat, sd)
    y_hat = quantize_dequantize(y)
    ctx = context_prediction(y_hat, sd)
    gp = _seq_conv(torch.cat((params, ctx), 1), sd, "entropy_parameters",
                   [(0, "conv", 1, 0), (2, "conv", 1, 0), (4, "conv", 1, 0)])
    scales, means = gp.chunk(2, 1)
    _, y_lik = gaussian_conditional_forward(y, scales, means)
    return {"y": y, "y_hat": y_hat, "x_hat": g_s(y_hat, sd, clamp=False), "lik_y": y_lik, "lik_z": z_lik,
            "scales": scales, "means": means}


def iframe_ar_code(x: Tensor, sd: SD):
    """compress() + decompress() of the I-frame model (priors.py:510-644)."""
    y = g_a(x, sd)
    z_hat = eb_quantize(iframe_hyper(y, sd), sd)
    t_hat, sym, idx = ar_scan(y, iframe_hyper_synthesis(z_hat, sd), sd, head="entropy_parameters")
    return {"y_hat": t_hat, "x_hat": g_s(t_hat, sd, clamp=True), "symbols": sym, "indexes": idx, "z_hat": z_hat}
