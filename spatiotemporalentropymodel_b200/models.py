"""Drop-in model classes for the STEM P-frame path, with the reference's names, constructor arguments,
state_dict keys and method signatures (compressai/models/spatiotemporalpriors.py, compressai/models/priors.py).

The nn.Module objects only *hold* parameters (so `load_state_dict` of reference checkpoints, `.to()`,
`.parameters()`, `.eval()` behave as in the reference); `forward` runs the CUDA engine
(spatiotemporalentropymodel_b200.engine).  Weights are repacked lazily on first use and re-packed after
`load_state_dict` / `.to()`.
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .engine import IFrameEntropyEngine, PFramePipeline, StemEngine, TransformsEngine, _require_cuda
from .entropy_models import EntropyBottleneck, GaussianConditional
from .synthetic import variant_flags

Tensor = torch.Tensor

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):  # noqa: A002 (reference signature)
    """spatiotemporalpriors.py:27-30"""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


# ------------------------------------------------------------------------------------------------------
# parameter-holding layers (state_dict compatible with compressai/layers)
# ------------------------------------------------------------------------------------------------------
class _LowerBoundBuf(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


class NonNegativeParametrizer(nn.Module):
    """ops/parametrizers.py:21-45 (buffers `pedestal`, `lower_bound.bound`)."""

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        pedestal = float(reparam_offset) ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = _LowerBoundBuf((float(minimum) + pedestal) ** 0.5)

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))


class GDN(nn.Module):
    """layers/gdn.py:22-67 parameter holder; the arithmetic is the GDN epilogue of the conv kernel."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))


class MaskedConv2d(nn.Conv2d):
    """layers/layers.py:21-47: mask buffer kept for state_dict parity; the kernel packs only the live taps."""

    def __init__(self, *args, mask_type: str = "A", **kwargs):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """models/utils.py:112-119"""
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    """models/utils.py:122-130"""
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def _resize_registered_buffers(module: nn.Module, prefix: str, names, state_dict) -> None:
    """models/utils.py:46-109 (`update_registered_buffers`, policy resize_if_empty): CDF buffers are empty
    until update(), so they are resized to the checkpoint's sizes before nn.Module.load_state_dict."""
    for name in names:
        key = f"{prefix}.{name}"
        if key not in state_dict:
            raise RuntimeError(f'Missing key "{key}" in state_dict')
        new = state_dict[key]
        cur = getattr(module, name)
        if cur.numel() == 0 or cur.size() != new.size():
            setattr(module, name, torch.empty(new.size(), dtype=cur.dtype, device=cur.device))


class CompressionModel(nn.Module):
    """priors.py:42-106: owns `entropy_bottleneck`, aux_loss, update, load_state_dict."""

    def __init__(self, entropy_bottleneck_channels: int, init_weights: bool = True):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        self._engine = None

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def update(self, force: bool = False) -> bool:
        updated = False
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                updated |= m.update(force=force)
        self._engine = None
        return updated

    def load_state_dict(self, state_dict, strict: bool = True):
        _resize_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                   ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        out = super().load_state_dict(state_dict, strict=strict)
        self._engine = None
        return out

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .float(): packed weights become stale
        self._engine = None
        return super()._apply(fn, *args, **kwargs)

    def _device(self) -> torch.device:
        return next(self.parameters()).device


# ------------------------------------------------------------------------------------------------------
# STEM P-frame entropy models
# ------------------------------------------------------------------------------------------------------
class _StemBase(CompressionModel):
    _variant = ""

    def __init__(self, entropy_bottleneck_channels: int = 256, in_channels: int = 192):
        super().__init__(entropy_bottleneck_channels=entropy_bottleneck_channels)
        has_tpm, has_spm, res = variant_flags(self._variant)
        self._flags = (has_tpm, has_spm, res)
        C, ebc = in_channels, entropy_bottleneck_channels
        if not (has_tpm or has_spm):
            ebc = 256  # WithoutSPMTPM hard-codes 256 (spatiotemporalpriors.py:43-58)
        if has_tpm:
            self.TPM = nn.Sequential(
                nn.Conv2d(C, 256, 5, padding=2, stride=1), nn.LeakyReLU(),
                nn.Conv2d(256, 320, 5, padding=2, stride=1), nn.LeakyReLU(),
                nn.Conv2d(320, C * 2, 5, padding=2, stride=1))
        self.HE = nn.Sequential(
            nn.Conv2d(C * 2, 256, 3, padding=1, stride=1), nn.LeakyReLU(),
            nn.Conv2d(256, 256, 5, padding=2, stride=2), nn.LeakyReLU(),
            nn.Conv2d(256, ebc, 5, padding=2, stride=2))
        self.HD = nn.Sequential(
            nn.ConvTranspose2d(ebc, 256, 5, padding=2, stride=2, output_padding=1), nn.LeakyReLU(),
            nn.ConvTranspose2d(256, 256, 5, padding=2, stride=2, output_padding=1), nn.LeakyReLU(),
            nn.Conv2d(256, C * 2, 3, padding=1, stride=1))
        if has_spm:
            self.context_prediction = MaskedConv2d(C, C * 2, kernel_size=5, padding=2, stride=1)
        n_prior = (1 + int(has_tpm) + int(has_spm)) * 2 * C
        self.EPM = nn.Sequential(
            nn.Conv2d(n_prior, 768, 1), nn.LeakyReLU(),
            nn.Conv2d(768, 576, 1), nn.LeakyReLU(),
            nn.Conv2d(576, C * 2, 1))
        self.gaussian_conditional = GaussianConditional(None)
        self.in_channels = in_channels

    # --- engine ---------------------------------------------------------------------------------------
    def engine(self) -> StemEngine:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("STEM models of spatiotemporalentropymodel_b200 run on CUDA only (no CPU fallback); "
                               "call .to('cuda') first")
        if self._engine is None:
            has_tpm, has_spm, res = self._flags
            gc = self.gaussian_conditional
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = StemEngine(
                sd, dev, has_tpm, has_spm, res, self.entropy_bottleneck.packed_params(), gc.scale_table,
                scale_bound=float(gc.lower_bound_scale.bound.item()),
                lik_bound=float(gc.likelihood_lower_bound.bound.item()))
        return self._engine

    # --- reference API --------------------------------------------------------------------------------
    def forward(self, y_cur: Tensor, y_conditioned: Tensor):
        """-> {"y_hat", "likelihoods": {"y", "z"}} (spatiotemporalpriors.py:561-585 and variants), eval mode."""
        if self.training:
            raise NotImplementedError("training-mode forward (noise quantisation + backward) is outside the "
                                      "inference hot path of this build; call .eval()")
        out = self.engine().forward_nchw(y_cur, y_conditioned)
        return {"y_hat": out["y_hat"], "likelihoods": out["likelihoods"]}

    def forward_with_indexes(self, y_cur: Tensor, y_conditioned: Tensor):
        """forward plus what compress() feeds the entropy coder: scale-table indexes (build_indexes,
        entropy_models.py:598-604) and symbols round(y - mu) (:148-150), from the same fused kernel."""
        return self.engine().forward_nchw(y_cur, y_conditioned, want_indexes=True)

    def compress(self, y_cur: Tensor, y_conditioned: Tensor):
        """Non-autoregressive variants (spatiotemporalpriors.py:86-96, :197-209):
        -> {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}. Symbols and CDF indexes come from the
        fused GaussianConditional kernel; the rANS coder is the C ABI's (byte-compatible with compressai.ans)."""
        from .engine import nchw_to_nhwc_f16, nhwc_f32_to_nchw
        eng = self.engine()
        _require_cuda(y_cur, y_conditioned)
        y_cur, y_conditioned = y_cur.contiguous().float(), y_conditioned.contiguous().float()
        B, C, h, w = y_cur.shape
        f16 = torch.float16
        y16 = nchw_to_nhwc_f16(y_cur, eng.ws.get("y16", (B, h, w, C), f16))
        cond16 = nchw_to_nhwc_f16(y_conditioned, eng.ws.get("cond16", (B, h, w, C), f16))
        z_nhwc = eng.hyper_latent(y16, cond16, B, h, w)
        z = nhwc_f32_to_nchw(z_nhwc, torch.empty((B, eng.zc, h // 4, w // 4), device=y_cur.device))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        zhat16 = nchw_to_nhwc_f16(z_hat.contiguous(), eng.ws.get("zhat16", (B, h // 4, w // 4, eng.zc), f16))
        if self._flags[1]:
            # autoregressive variants (spatiotemporalpriors.py:588-678, _Res :871-961): wavefront kernel over the
            # latent grid, symbols / indexes in the reference's raster (h, w, c) stream order
            priors = eng.static_priors(zhat16, cond16, B, h, w)
            target = (y_cur - y_conditioned) if self._flags[2] else y_cur
            target = target.permute(0, 2, 3, 1).contiguous()
            _, sym, idx, _ = eng.ar_head().encode(target, priors, eng.scale_table)
            y_strings = self.gaussian_conditional.compress_symbols(sym, idx)
            return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}
        params = eng.params_from_zhat(zhat16, cond16, None, B, h, w)
        idx = torch.empty(y_cur.shape, dtype=torch.int32, device=y_cur.device)
        sym = torch.empty(y_cur.shape, dtype=torch.int32, device=y_cur.device)
        eng.gaussian_conditional(y_cur, True, None, params, B, h, w, None, None, idx, sym)
        y_strings = self.gaussian_conditional.compress_symbols(sym, idx)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape, y_conditioned: Tensor):
        """Non-autoregressive variants (spatiotemporalpriors.py:99-111, :212-225). Returns a dict with "y_hat" and
        "entropy_params" (what stem/evalSTEM.py:120,152 reads; the reference returns the bare tensor)."""
        assert isinstance(strings, list) and len(strings) == 2
        from .engine import nchw_to_nhwc_f16
        eng = self.engine()
        _require_cuda(y_conditioned)
        y_conditioned = y_conditioned.contiguous().float()
        B, C, h, w = y_conditioned.shape
        f16 = torch.float16
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape).to(y_conditioned.device)
        cond16 = nchw_to_nhwc_f16(y_conditioned, eng.ws.get("cond16", (B, h, w, C), f16))
        zhat16 = nchw_to_nhwc_f16(z_hat.contiguous(), eng.ws.get("zhat16", (B, h // 4, w // 4, eng.zc), f16))
        if self._flags[1]:
            # autoregressive variants (spatiotemporalpriors.py:681-768, _Res :964-1055): raster-order kernel with the
            # rANS decoder inside; the decoder repeats the encoder's arithmetic, so (idx, mu) match bit for bit
            gc = self.gaussian_conditional
            priors = eng.static_priors(zhat16, cond16, B, h, w)
            t_hat, params = eng.ar_head().decode(strings[0], priors, B, h, w, eng.scale_table, gc.quantized_cdf,
                                                 gc.cdf_length, gc.offset)
            y_hat = t_hat.permute(0, 3, 1, 2).contiguous()
            if self._flags[2]:
                y_hat = y_hat + y_conditioned
            gp = params.permute(0, 3, 1, 2)
            return {"y_hat": y_hat, "entropy_params": {"scales_hat": gp[:, :C].contiguous(),
                                                       "means_hat": gp[:, C:].contiguous()}}
        params = eng.params_from_zhat(zhat16, cond16, None, B, h, w)
        idx = torch.empty(y_conditioned.shape, dtype=torch.int32, device=y_conditioned.device)
        eng.gaussian_conditional(torch.zeros_like(y_conditioned), True, None, params, B, h, w, None, None, idx, None)
        gp = params.permute(0, 3, 1, 2)
        scales_hat, means_hat = gp[:, :C].contiguous(), gp[:, C:].contiguous()
        y_hat = self.gaussian_conditional.decompress(strings[0], idx, means=means_hat)
        return {"y_hat": y_hat, "entropy_params": {"scales_hat": scales_hat, "means_hat": means_hat}}

    def load_state_dict(self, state_dict, strict: bool = True):
        _resize_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                   ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    def update(self, scale_table=None, force: bool = False) -> bool:
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated


class SpatioTemporalPriorModelWithoutSPMTPM(_StemBase):
    """spatiotemporalpriors.py:33-129"""
    _variant = "SpatioTemporalPriorModelWithoutSPMTPM"


class SpatioTemporalPriorModelWithoutSPM(_StemBase):
    """spatiotemporalpriors.py:132-243"""
    _variant = "SpatioTemporalPriorModelWithoutSPM"


class SpatioTemporalPriorModelWithoutTPM(_StemBase):
    """spatiotemporalpriors.py:246-505"""
    _variant = "SpatioTemporalPriorModelWithoutTPM"


class SpatioTemporalPriorModel(_StemBase):
    """spatiotemporalpriors.py:508-788"""
    _variant = "SpatioTemporalPriorModel"


class SpatioTemporalPriorModel_Res(_StemBase):
    """spatiotemporalpriors.py:791-1072"""
    _variant = "SpatioTemporalPriorModel_Res"


# ------------------------------------------------------------------------------------------------------
# I-frame model shell: only g_a (getY) and g_s (getX) are on the P-frame path
# ------------------------------------------------------------------------------------------------------
class JointAutoregressiveHierarchicalPriors(CompressionModel):
    """priors.py:406-694 parameter layout (so mbt2018 checkpoints load); getY / getX run on the CUDA engine.
    The I-frame codec itself (forward/compress/decompress, :477-684) is a 'next' row."""

    def __init__(self, N: int = 192, M: int = 192, **kwargs):
        super().__init__(entropy_bottleneck_channels=N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5))
        self.h_s = nn.Sequential(deconv(N, M, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 deconv(M, M * 3 // 2, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))
        self.gaussian_conditional = GaussianConditional(None)
        self.entropy_parameters = nn.Sequential(
            nn.Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
            nn.Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.N, self.M = int(N), int(M)
        self._entropy_engine = None
        self._entropy_engine_src = None

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def engine(self) -> TransformsEngine:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("g_a / g_s of spatiotemporalentropymodel_b200 run on CUDA only (no CPU fallback)")
        if self._engine is None:
            self._engine = TransformsEngine(dict(self.state_dict()), dev)
        return self._engine

    def getY(self, x: Tensor):
        """priors.py:686-694: (y, y_quantized) with y_quantized = gaussian_conditional.quantize(y, "noise") =
        y + U(-1/2, 1/2) (entropy_models.py:128-131) - the reference adds the noise even in eval mode; evalSTEM ignores
        the second output (:113)."""
        from .engine import nhwc_f32_to_nchw
        eng = self.engine()
        y_nhwc, h, w = eng.analysis(x if x.dtype == torch.uint8 else x.float())
        y = torch.empty((x.shape[0], self.M, h, w), dtype=torch.float32, device=x.device)
        nhwc_f32_to_nchw(y_nhwc, y)
        return y, y + torch.empty_like(y).uniform_(-0.5, 0.5)

    def getX(self, y_hat: Tensor) -> Tensor:
        """priors.py:397-402: g_s(y_hat).clamp_(0, 1)"""
        from .engine import nchw_to_nhwc_f16
        _require_cuda(y_hat)
        eng = self.engine()
        B, C, h, w = y_hat.shape
        y16 = nchw_to_nhwc_f16(y_hat.contiguous().float(), eng.ws.get("getx_in", (B, h, w, C), torch.float16))
        out = torch.empty((B, 3, 16 * h, 16 * w), dtype=torch.float32, device=y_hat.device)
        return eng.synthesis(y16, out=out)

    def entropy_engine(self) -> IFrameEntropyEngine:
        """h_a / h_s / context_prediction / entropy_parameters + EntropyBottleneck on the CUDA kernels."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("the I-frame model of spatiotemporalentropymodel_b200 runs on CUDA only")
        if self.M != 192:
            # entropy_parameters widths 10M/3 and 8M/3 (1066 / 853 for M = 320) are not multiples of 64, which the
            # tensor-core tiles and the autoregressive kernel's 64-CTA row split need
            raise NotImplementedError(
                f"the I-frame entropy model (forward / compress / decompress) is built for M = 192 (mbt2018 quality "
                f"1-4); this model has M = {self.M} (quality 5-8). getY / getX work for every quality.")
        tr = self.engine()
        if self._entropy_engine is None or self._entropy_engine_src is not tr:  # rebuilt after load / update / .to()
            self._entropy_engine_src = tr
            gc = self.gaussian_conditional
            self._entropy_engine = IFrameEntropyEngine(
                dict(self.state_dict()), dev, self.entropy_bottleneck.packed_params(), gc.scale_table,
                scale_bound=float(gc.lower_bound_scale.bound.item()),
                lik_bound=float(gc.likelihood_lower_bound.bound.item()))
        return self._entropy_engine

    def _latents(self, x: Tensor):
        """g_a(x) as NHWC fp32 + fp16 copies; x must have sides that are multiples of 64 (downsampling_factor)."""
        _require_cuda(x)
        if x.shape[2] % 64 or x.shape[3] % 64:
            raise ValueError("the I-frame model needs frame sides that are multiples of 64 (priors.py:473-475)")
        eng, ee = self.engine(), self.entropy_engine()
        y32, h, w = eng.analysis(x.float())
        B = x.shape[0]
        lib = _lib.load()
        y16 = ee.ws.get("y16", (B, h, w, self.M), torch.float16)
        yq16 = ee.ws.get("yq16", (B, h, w, self.M), torch.float16)
        _lib.check(lib.stemb200_latent_stage(y32.data_ptr(), None, y16.data_ptr(), yq16.data_ptr(), None, y32.numel(),
                                             torch.cuda.current_stream().cuda_stream), "latent_stage")
        return y32, y16, yq16, B, h, w

    def forward(self, x):
        """priors.py:477-508 (eval mode): -> {"y", "y_hat", "x_hat", "likelihoods", "entropy_params"};
        y_hat = round(y), likelihoods of y at (sigma, mu) from the context of round(y), x_hat = g_s(y_hat)."""
        if self.training:
            raise NotImplementedError("training-mode forward is outside the inference path of this build; call .eval()")
        from .engine import nhwc_f32_to_nchw
        eng, ee = self.engine(), self.entropy_engine()
        y32, y16, yq16, B, h, w = self._latents(x)
        dev = x.device
        zc, h4, w4 = ee.zc, h // 4, w // 4
        z_lik = torch.empty((B, zc, h4, w4), dtype=torch.float32, device=dev)
        bits = torch.zeros((2, B), dtype=torch.float64, device=dev)
        params = ee.gaussian_params(y16, None, yq16, B, h, w, None, z_lik, bits[1])
        y_hat = torch.empty((B, self.M, h, w), dtype=torch.float32, device=dev)
        y_lik = torch.empty_like(y_hat)
        ee.gaussian_conditional(y32, False, None, params, B, h, w, y_hat, y_lik, bits=bits[0])
        y = nhwc_f32_to_nchw(y32, torch.empty_like(y_hat))
        x_hat = torch.empty((B, 3, 16 * h, 16 * w), dtype=torch.float32, device=dev)
        eng.synthesis(yq16, out=x_hat, clamp=False)
        gp = params.permute(0, 3, 1, 2)
        return {"y": y, "y_hat": y_hat, "x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik},
                "entropy_params": {"scales_hat": gp[:, :self.M].contiguous(), "means_hat": gp[:, self.M:].contiguous()}}

    def compress(self, x):
        """priors.py:510-600 -> {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}; the per-position scan
        (:556-600) runs as wavefronts on the GPU (csrc/ar_codec.cu)."""
        from .engine import nchw_to_nhwc_f16, nhwc_f32_to_nchw
        ee = self.entropy_engine()
        y32, y16, _, B, h, w = self._latents(x)
        z_nhwc = ee.hyper_latent(y16, None, B, h, w)
        z = nhwc_f32_to_nchw(z_nhwc, torch.empty((B, ee.zc, h // 4, w // 4), device=x.device))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        zhat16 = nchw_to_nhwc_f16(z_hat.contiguous(), ee.ws.get("zhat16", (B, h // 4, w // 4, ee.zc), torch.float16))
        priors = ee.static_priors(zhat16, None, B, h, w)
        _, sym, idx, _ = ee.ar_head().encode(y32, priors, ee.scale_table)
        y_strings = self.gaussian_conditional.compress_symbols(sym, idx)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape):
        """priors.py:602-644 -> {"x_hat" (clamped), "y_hat"}; raster-order GPU scan with the rANS decoder inside."""
        from .engine import nchw_to_nhwc_f16
        assert isinstance(strings, list) and len(strings) == 2
        eng, ee = self.engine(), self.entropy_engine()
        dev = self._device()
        gc = self.gaussian_conditional
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape).to(dev)
        B, _, h4, w4 = z_hat.shape
        h, w = 4 * h4, 4 * w4
        zhat16 = nchw_to_nhwc_f16(z_hat.contiguous(), ee.ws.get("zhat16", (B, h4, w4, ee.zc), torch.float16))
        priors = ee.static_priors(zhat16, None, B, h, w)
        t_hat, _ = ee.ar_head().decode(strings[0], priors, B, h, w, ee.scale_table, gc.quantized_cdf, gc.cdf_length,
                                       gc.offset)
        y_hat = t_hat.permute(0, 3, 1, 2).contiguous()
        return {"x_hat": self.getX(y_hat), "y_hat": y_hat}

    def load_state_dict(self, state_dict, strict: bool = True):
        _resize_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                   ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    def update(self, scale_table=None, force: bool = False) -> bool:
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated


# zoo: (N, M) per quality for mbt2018 (zoo/image.py:162-171)
_MBT2018_CFG = {1: (192, 192), 2: (192, 192), 3: (192, 192), 4: (192, 192), 5: (192, 320), 6: (192, 320),
                7: (192, 320), 8: (192, 320)}


def mbt2018(quality: int, metric: str = "mse", pretrained: bool = False, progress: bool = True, **kwargs):
    """zoo/image.py:_load_model for "mbt2018" (pretrained download needs network: unsupported here)."""
    if pretrained:
        raise RuntimeError("pretrained zoo weights need network access; load a checkpoint with load_state_dict")
    if quality not in _MBT2018_CFG:
        raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
    return JointAutoregressiveHierarchicalPriors(*_MBT2018_CFG[quality], **kwargs)


models = {"mbt2018": mbt2018}


def make_pipeline(iframe_model: JointAutoregressiveHierarchicalPriors, stem_model: _StemBase) -> PFramePipeline:
    """GOP-batched P-frame pipeline (pad -> g_a -> STEM -> g_s -> bit / distortion sums)."""
    return PFramePipeline(iframe_model.engine(), stem_model.engine())
