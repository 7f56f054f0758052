"""Host-side execution engine: weight repacking, workspace, and the kernel chains of the P-frame path.

All arithmetic happens in libstemb200.so (hand-written sm_100a kernels); PyTorch is used for device memory,
streams and parameter storage only.  Nothing here falls back to torch ops for compute: a non-CUDA tensor or a
missing library raises.

Kernel chains (reference lines in parentheses):
  analysis   g_a  (priors.py:421-429)      frame canvas -> [row-taps conv+GDN] -> [conv s2+GDN] x2 -> conv s2 (fp32 y)
  STEM            (spatiotemporalpriors.py:561-585 and variants) {HE -> EB -> HD} || {TPM, ctx} -> EPM.0, EPM.2 ->
                  [EPM.4 + GaussianConditional]
  synthesis  g_s  (priors.py:431-439,397-402)  [deconv+IGDN] x2 -> [deconv+IGDN+last-layer GEMM] -> col2im (clamp, MSE),
                  on a side stream next to the entropy model where y_hat does not depend on it
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import DT_F16, DT_F32, EPI_ADD, EPI_LINEAR, EPI_SFT, ArDesc, ConvDesc

Tensor = torch.Tensor
MASK_A_5x5 = 0x00000FFF  # taps (r, s) with r < 2, or r == 2 and s < 2 (layers/layers.py:39-42)
SQ_SCALE = 0.125         # x^2 is carried as (x/8)^2 in fp16 (|x| < 2047 stays finite)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spatiotemporalentropymodel_b200 runs on CUDA (sm_100a) tensors only; "
                               "there is no CPU fallback")


def fuse_gc_enabled() -> bool:
    """entropy_parameters' last layer and GaussianConditional run as one kernel in the GOP pipeline (sigma, mu never in
    HBM; BASELINE north_star item 2). On by default: bit-identical to the two separate kernels, 155 us instead of
    65 + 119 us per 11 x 1080p latents and 226 MB less DRAM traffic (profiles/r02_ncu_fused_gc.txt). STEMB200_FUSE_GC=0
    selects the separate kernels (which also leave sigma | mu in the workspace buffer "gparams")."""
    return os.environ.get("STEMB200_FUSE_GC", "1") != "0"


def overlap_level() -> int:
    """STEMB200_OVERLAP: 0 = one stream; 1 = the temporal-prior / context chain runs beside the hyper-prior chain;
    2 (default) = in addition the synthesis transform of a GOP runs beside its entropy model (SPM variants, where
    y_hat = round(y) is known before the entropy parameters are)."""
    try:
        return int(os.environ.get("STEMB200_OVERLAP", "2"))
    except ValueError:
        return 2


class SideStream:
    """Fork / join of a second CUDA stream around a block of launches.  The launches under the `with` go to the side
    stream, which first waits for everything issued so far on the current stream; `join()` makes the current stream
    wait for them.  Under stream capture the two branches become parallel branches of the CUDA graph.  Kernels of the
    path are persistent (one CTA per SM), so what this buys is the tails: SMs a small launch leaves idle run the other
    branch's tiles."""

    def __init__(self, device: torch.device):
        self.stream = torch.cuda.Stream(device=device)
        self._joined = True

    def __enter__(self):
        self.main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.stream.wait_event(ev)
        self._ctx = torch.cuda.stream(self.stream)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._done = torch.cuda.Event()
        self._done.record(self.stream)
        self._joined = False
        return self._ctx.__exit__(*exc)

    def join(self):
        if not self._joined:
            torch.cuda.current_stream().wait_event(self._done)
            self._joined = True


class Workspace:
    """Named, shape-keyed device buffers that persist across calls (so CUDA graphs can capture the chain)."""

    def __init__(self, device: torch.device):
        self.device = device
        self._bufs: Dict[str, Tensor] = {}

    def get(self, name: str, shape: Sequence[int], dtype: torch.dtype) -> Tensor:
        t = self._bufs.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._bufs.values())


class ConvOp:
    """One packed convolution: geometry + repacked fp16 weight matrix + fp32 bias on the device."""

    def __init__(self, weight: Tensor, bias: Tensor, *, c_in: Sequence[int], c_out: int, k: int, stride: int = 1,
                 transposed: bool = False, tap_mask: int = 0, slope: float = 1.0, out_dtype: int = DT_F16,
                 direct_store: bool = False, gdn: Optional[Tuple[Tensor, Tensor, bool]] = None,
                 alg_flops_per_out_pixel: Optional[float] = None, epilogue: int = EPI_LINEAR,
                 row_taps: bool = False):
        """gdn = (beta', gamma', inverse) with the re-parametrised beta (C,) / gamma (C, C): the layer is followed
        by GDN / IGDN and both run in one kernel (stemb200_conv2d_gdn_fwd)."""
        _require_cuda(weight, bias)
        self.lib = _lib.load()
        self.c_in = list(c_in)
        self.c_out = c_out
        self.k, self.stride, self.transposed = k, stride, transposed
        self.out_dtype = out_dtype
        d = ConvDesc()
        d.batch, d.h_in, d.w_in = 1, 16, 16
        d.n_src = len(c_in)
        for i, c in enumerate(c_in):
            d.c_in[i] = c
        d.c_out = c_out
        d.kh = d.kw = k
        d.stride = stride
        d.transposed = int(transposed)
        d.tap_mask = tap_mask
        d.epilogue = epilogue
        d.lrelu_slope = slope
        d.out_dtype = out_dtype
        d.sq_scale = SQ_SCALE
        d.tile_h = d.tile_w = 0
        d.direct_store = int(direct_store)
        d.row_taps = int(row_taps)
        self.desc = d
        K = self.lib.stemb200_conv2d_packed_k(C.byref(d))
        if K <= 0:
            _lib.check(int(K), "conv2d_packed_k")
        w32 = weight.detach().to(torch.float32).contiguous()
        self.packed = torch.empty((c_out, K), dtype=torch.float16, device=weight.device)
        _lib.check(self.lib.stemb200_conv2d_pack_weight(C.byref(d), w32.data_ptr(), self.packed.data_ptr(),
                                                        _stream()), "conv2d_pack_weight")
        self.bias = bias.detach().to(torch.float32).contiguous()
        # algorithmic FLOPs per output pixel: 2*C_in*C_out*taps (deconv: per *input* pixel, i.e. /4
        # per output pixel), masked taps not counted, + 2*C*C for a fused GDN
        ntaps = bin(tap_mask).count("1") if tap_mask else k * k
        alg = 2.0 * sum(c_in) * c_out * ntaps / (4.0 if transposed else 1.0)
        if gdn is not None:
            alg += 2.0 * c_out * c_out
        self.alg_flops_per_out_pixel = alg if alg_flops_per_out_pixel is None else float(alg_flops_per_out_pixel)
        self.gdn = None
        if gdn is not None:
            beta, gamma, inverse = gdn
            gd = ConvDesc()
            gd.batch, gd.h_in, gd.w_in, gd.n_src = 1, 8, 8, 1
            gd.c_in[0] = c_out
            gd.c_out, gd.kh, gd.kw, gd.stride = c_out, 1, 1, 1
            gd.lrelu_slope, gd.out_dtype, gd.sq_scale = 1.0, DT_F16, SQ_SCALE
            g32 = gamma.detach().to(weight.device, torch.float32).reshape(c_out, c_out, 1, 1).contiguous()
            self.gamma_packed = torch.empty((c_out, c_out), dtype=torch.float16, device=weight.device)
            _lib.check(self.lib.stemb200_conv2d_pack_weight(C.byref(gd), g32.data_ptr(), self.gamma_packed.data_ptr(),
                                                            _stream()), "conv2d_pack_weight(gamma)")
            self.beta = beta.detach().to(weight.device, torch.float32).contiguous()
            self.gdn = bool(inverse)

    def call_last(self, inputs: Sequence[Tensor], batch: int, h: int, w: int, w6_packed: Tensor, col: Tensor,
                  act: Optional[Tensor] = None) -> Tensor:
        """deconv + IGDN with the final deconv(N, 3) GEMM fused behind it (stemb200_conv2d_gdn_last_fwd):
        col = (B, 2h, 2w, 96) fp16 tap contributions; act = the layer's own activation or None (not written)."""
        if self.gdn is not True:
            raise ValueError("call_last needs a deconv + IGDN layer")
        d = self.desc
        d.batch, d.h_in, d.w_in = batch, h, w
        arr = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        _lib.check(self.lib.stemb200_conv2d_gdn_last_fwd(
            C.byref(d), arr, self.packed.data_ptr(), self.bias.data_ptr(), self.gamma_packed.data_ptr(),
            self.beta.data_ptr(), w6_packed.data_ptr(), col.data_ptr(), _ptr(act), _stream()), "conv2d_gdn_last_fwd")
        return col

    def alg_flops(self, batch: int, h: int, w: int) -> float:
        ho, wo = self.out_hw(h, w)
        return self.alg_flops_per_out_pixel * batch * ho * wo

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        if self.transposed:
            return 2 * h, 2 * w
        if self.stride == 2:
            p = self.k // 2
            return (h + 2 * p - self.k) // 2 + 1, (w + 2 * p - self.k) // 2 + 1
        return h, w

    def __call__(self, inputs: Sequence[Tensor], batch: int, h: int, w: int, out: Tensor,
                 aux: Optional[Tensor] = None) -> Tensor:
        d = self.desc
        d.batch, d.h_in, d.w_in = batch, h, w
        arr = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        if self.gdn is None:
            _lib.check(self.lib.stemb200_conv2d_fwd(C.byref(d), arr, self.packed.data_ptr(), self.bias.data_ptr(),
                                                    _ptr(aux), out.data_ptr(), _stream()), "conv2d_fwd")
        else:
            _lib.check(self.lib.stemb200_conv2d_gdn_fwd(C.byref(d), arr, self.packed.data_ptr(), self.bias.data_ptr(),
                                                        self.gamma_packed.data_ptr(), self.beta.data_ptr(),
                                                        int(self.gdn), out.data_ptr(), _stream()), "conv2d_gdn_fwd")
        return out


class FirstLayerOp:
    """g_a.0 + GDN with resident weights (csrc/conv_first.cu, stemb200_conv_first_gdn_fwd): the frame is staged as an
    NHWC4 canvas and W0 is packed as [192][5 rows][8 taps x 4 channels]."""

    def __init__(self, weight: Tensor, bias: Tensor, beta: Tensor, gamma: Tensor):
        _require_cuda(weight, bias)
        N = weight.shape[0]
        if tuple(weight.shape) != (192, 3, 5, 5):
            raise ValueError("FirstLayerOp is built for conv(3, 192, k5, s2)")
        dev = weight.device
        self.lib = _lib.load()
        w = torch.zeros((N, 5, 8, 4), dtype=torch.float32, device=dev)
        w[:, :, :5, :3] = weight.detach().float().permute(0, 2, 3, 1)   # [o][r][s][ch]
        self.packed = w.reshape(N, 160).to(torch.float16).contiguous()
        self.bias = bias.detach().float().contiguous()
        self.beta = beta.detach().to(dev, torch.float32).contiguous()
        self.gamma_packed = gamma.detach().to(dev, torch.float32).to(torch.float16).contiguous()  # [c_out][c_in], K-major
        self.c_out = N
        self.alg_flops_per_out_pixel = 2.0 * 75 * N + 2.0 * N * N
        self.gdn = False

    def alg_flops(self, batch: int, h: int, w: int) -> float:
        return self.alg_flops_per_out_pixel * batch * (h // 2) * (w // 2)

    def out_hw(self, h: int, w: int) -> Tuple[int, int]:
        return h // 2, w // 2

    def __call__(self, inputs: Sequence[Tensor], batch: int, h: int, w: int, out: Tensor,
                 aux: Optional[Tensor] = None) -> Tensor:
        _lib.check(self.lib.stemb200_conv_first_gdn_fwd(
            inputs[0].data_ptr(), batch, h, w, 2, self.packed.data_ptr(), self.bias.data_ptr(),
            self.gamma_packed.data_ptr(), self.beta.data_ptr(), SQ_SCALE, out.data_ptr(), _stream()),
            "conv_first_gdn_fwd")
        return out


def sft_op(w_gamma: Tensor, b_gamma: Tensor, w_beta: Tensor, b_beta: Tensor, c_in: int, slope: float = 1.0) -> ConvOp:
    """SFT.mlp_gamma / mlp_beta (stem_utils.py:33-34) as ONE conv whose epilogue applies x*(1+gamma)+beta
    (:41): output rows are interleaved per 64 channels as [gamma(64) | beta(64)], and the "+1" goes into the
    gamma bias. Returns a ConvOp to be called with aux = x; its output has x's channel count."""
    C = w_gamma.shape[0]
    if C % 64:
        raise ValueError("SFT channel count must be a multiple of 64")
    k = w_gamma.shape[-1]
    ws, bs = [], []
    for g0 in range(0, C, 64):
        ws += [w_gamma[g0:g0 + 64], w_beta[g0:g0 + 64]]
        bs += [b_gamma[g0:g0 + 64] + 1.0, b_beta[g0:g0 + 64]]
    op = ConvOp(torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous(), c_in=[c_in], c_out=2 * C, k=k,
                slope=slope, epilogue=EPI_SFT,
                alg_flops_per_out_pixel=2.0 * c_in * 2 * C * k * k)
    op.c_out_real = C
    return op


def avgpool_nhwc(x: Tensor, out: Tensor, factor: int) -> Tensor:
    n, ho, wo, c = out.shape
    _lib.check(_lib.load().stemb200_avgpool_nhwc_f16(x.data_ptr(), out.data_ptr(), n, ho, wo, c, factor, _stream()),
               "avgpool_nhwc_f16")
    return out


# ------------------------------------------------------------------------------------------------------
# thin wrappers of the elementwise entry points
# ------------------------------------------------------------------------------------------------------
def nchw_to_nhwc_f16(x: Tensor, out: Tensor, sub: Optional[Tensor] = None, round_first: bool = False) -> Tensor:
    n, c, h, w = x.shape
    _lib.check(_lib.load().stemb200_nchw_f32_to_nhwc_f16(x.data_ptr(), _ptr(sub), out.data_ptr(), n, c, h, w,
                                                         int(round_first), _stream()), "nchw_to_nhwc_f16")
    return out


def nhwc_f16_to_nchw(x: Tensor, out: Tensor) -> Tensor:
    n, h, w, c = x.shape
    _lib.check(_lib.load().stemb200_nhwc_f16_to_nchw_f32(x.data_ptr(), out.data_ptr(), n, c, h, w, _stream()),
               "nhwc_f16_to_nchw_f32")
    return out


def nhwc_f32_to_nchw(x: Tensor, out: Tensor) -> Tensor:
    n, h, w, c = x.shape
    _lib.check(_lib.load().stemb200_nhwc_f32_to_nchw_f32(x.data_ptr(), out.data_ptr(), n, c, h, w, _stream()),
               "nhwc_f32_to_nchw_f32")
    return out


def gaussian_conditional_flat(y: Tensor, scales: Tensor, means: Optional[Tensor], table: Optional[Tensor],
                              scale_bound: float, lik_bound: float, want_yhat=True, want_lik=True,
                              want_idx=False, want_sym=False, want_bits=False):
    """Flat-array GaussianConditional (entropy_models.py:588-604) on CUDA fp32 tensors of equal shape."""
    _require_cuda(y, scales, means)
    y = y.contiguous()
    scales = scales.contiguous()
    means = means.contiguous() if means is not None else None
    dev = y.device
    y_hat = torch.empty_like(y) if want_yhat else None
    lik = torch.empty_like(y) if want_lik else None
    idx = torch.empty(y.shape, dtype=torch.int32, device=dev) if want_idx else None
    sym = torch.empty(y.shape, dtype=torch.int32, device=dev) if want_sym else None
    bits = torch.zeros(1, dtype=torch.float64, device=dev) if want_bits else None
    n_scales = 0 if table is None else table.numel()
    _lib.check(_lib.load().stemb200_gaussian_conditional_flat(
        y.data_ptr(), scales.data_ptr(), _ptr(means), y.numel(), _ptr(table), n_scales, scale_bound, lik_bound,
        _ptr(y_hat), _ptr(lik), _ptr(idx), _ptr(sym), _ptr(bits), _stream()), "gaussian_conditional_flat")
    return y_hat, lik, idx, sym, bits


# ------------------------------------------------------------------------------------------------------
# autoregressive coding head (context_prediction + the three 1x1 layers) for compress() / decompress()
# ------------------------------------------------------------------------------------------------------
AR_CTAS = 64
AR_TAPS = [(r, s_) for r in range(3) for s_ in range(5) if r < 2 or s_ < 2]  # mask 'A' order (layers.py:39-42)


class ArHead:
    """spatiotemporalpriors.py:633-678 / :729-768 and priors.py:556-600 / :651-684 on the persistent AR kernel
    (csrc/ar_codec.cu). The prior columns of the first 1x1 layer are a batched GEMM ("e0", tcgen05 conv kernel);
    context_prediction and the rest of the head run per latent position inside the kernel, fp32."""

    def __init__(self, w_ctx: Tensor, b_ctx: Tensor, layers: Sequence[Tuple[Tensor, Tensor]], static_c: Sequence[int],
                 device: torch.device, slope: float = 0.01, scale_bound: float = 0.11):
        f = lambda t: t.detach().to(device, torch.float32)
        w_ctx, b_ctx = f(w_ctx), f(b_ctx)
        (w0, b0), (w1, b1), (w2, b2) = [(f(w).flatten(1), f(b)) for w, b in layers]
        C2, C = w_ctx.shape[0], w_ctx.shape[1]
        L1, L2 = w0.shape[0], w1.shape[0]
        n_static = sum(static_c)
        if w0.shape[1] != n_static + C2 or w2.shape[0] != C2:
            raise ValueError("ArHead: inconsistent layer shapes")
        self.C, self.L1, self.L2, self.slope = C, L1, L2, float(slope)
        self.scale_bound = float(scale_bound)
        self.device = device
        self.lib = _lib.load()
        # e0 = W0[:, :n_static] . priors + b0, no activation, fp32 (the context columns are added per position)
        self.e0_conv = ConvOp(w0[:, :n_static].reshape(L1, n_static, 1, 1).contiguous(), b0, c_in=list(static_c),
                              c_out=L1, k=1, out_dtype=DT_F32)
        rc, r1, r2, rg = C2 // AR_CTAS, L1 // AR_CTAS, L2 // AR_CTAS, C // AR_CTAS
        wc = torch.stack([w_ctx[:, :, r, s_] for r, s_ in AR_TAPS], dim=1).reshape(C2, len(AR_TAPS) * C)
        blocks = torch.cat([
            wc.reshape(AR_CTAS, rc * wc.shape[1]), b_ctx.reshape(AR_CTAS, rc),
            w0[:, n_static:].reshape(AR_CTAS, r1 * C2),
            w1.reshape(AR_CTAS, r2 * L1), b1.reshape(AR_CTAS, r2),
            w2[:C].reshape(AR_CTAS, rg * L2), w2[C:].reshape(AR_CTAS, rg * L2),
            b2[:C].reshape(AR_CTAS, rg), b2[C:].reshape(AR_CTAS, rg)], dim=1).contiguous()
        self.packed = blocks
        self._ws: Dict[Tuple[int, int, int], Tensor] = {}

    def _desc(self, B: int, h: int, w: int, n_scales: int) -> ArDesc:
        d = ArDesc()
        d.batch, d.h, d.w, d.c, d.l1, d.l2, d.slope, d.n_scales = B, h, w, self.C, self.L1, self.L2, self.slope, n_scales
        d.scale_bound = self.scale_bound
        want = self.lib.stemb200_ar_packed_floats(C.byref(d))
        if want != self.packed.numel():
            raise _lib.StemLibError(f"AR weight block mismatch: packed {self.packed.numel()} floats, kernel wants {want}")
        return d

    def _workspace(self, d: ArDesc) -> Tensor:
        n = int(self.lib.stemb200_ar_workspace_bytes(C.byref(d)))
        if n <= 0:
            _lib.check(n, "ar_workspace_bytes")
        key = (d.batch, d.h, d.w)
        if key not in self._ws:
            self._ws[key] = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._ws[key]

    def _check_completed(self, ws: Tensor, what: str) -> None:
        """The persistent grid aborts when one of its grid barriers times out and then leaves its outputs undefined:
        words [1] (abort) and [2] (completed) of the workspace say which (include/stemb200.h)."""
        flags = ws[:12].view(torch.int32).cpu()
        if int(flags[1]) != 0 or int(flags[2]) != 1:
            raise _lib.StemLibError(f"{what}: the autoregressive kernel aborted (grid barrier watchdog; abort flag "
                                    f"{int(flags[1])}, completion flag {int(flags[2])}); its outputs are invalid")

    def static_part(self, priors: Sequence[Tensor], B: int, h: int, w: int) -> Tensor:
        out = torch.empty((B, h, w, self.L1), dtype=torch.float32, device=self.device)
        return self.e0_conv(list(priors), B, h, w, out)

    def encode(self, target_nhwc: Tensor, priors: Sequence[Tensor], table: Tensor, e0: Optional[Tensor] = None):
        """target (B, h, w, C) fp32 NHWC -> t_hat, symbols, indexes (stream order), params (sigma | mu).
        e0 (B, h, w, L1) fp32: the prior part of the first head layer computed elsewhere (tests feed the oracle's fp32
        value to separate the fp16 tensor-core priors from the head's own summation order); default: static_part."""
        B, h, w, _ = target_nhwc.shape
        d = self._desc(B, h, w, table.numel())
        if e0 is None:
            e0 = self.static_part(priors, B, h, w)
        elif tuple(e0.shape) != (B, h, w, self.L1) or e0.dtype != torch.float32 or not e0.is_contiguous():
            raise ValueError("e0 must be a contiguous (B, h, w, L1) fp32 tensor")
        dev = self.device
        t_hat = torch.empty_like(target_nhwc)
        sym = torch.empty(target_nhwc.shape, dtype=torch.int32, device=dev)
        idx = torch.empty(target_nhwc.shape, dtype=torch.int32, device=dev)
        params = torch.empty((B, h, w, 2 * self.C), dtype=torch.float32, device=dev)
        ws = self._workspace(d)
        _lib.check(self.lib.stemb200_ar_encode(C.byref(d), self.packed.data_ptr(), e0.data_ptr(), target_nhwc.data_ptr(),
                                               table.data_ptr(), t_hat.data_ptr(), sym.data_ptr(), idx.data_ptr(),
                                               params.data_ptr(), ws.data_ptr(), _stream()), "ar_encode")
        self._check_completed(ws, "ar_encode")
        return t_hat, sym, idx, params

    def decode(self, strings: Sequence[bytes], priors: Sequence[Tensor], B: int, h: int, w: int, table: Tensor,
               cdf: Tensor, cdf_length: Tensor, offset: Tensor):
        """-> t_hat (B, h, w, C) fp32 NHWC, params (sigma | mu). Raises on a corrupt stream."""
        import numpy as np
        if len(strings) != B:
            raise ValueError("one string per batch element expected")
        d = self._desc(B, h, w, table.numel())
        e0 = self.static_part(priors, B, h, w)
        dev = self.device
        offs, lens, pos = [], [], 0
        for s_ in strings:
            if len(s_) % 4 or len(s_) < 8:
                raise ValueError("rANS strings are whole 32-bit words (>= 2)")
            offs.append(pos)
            lens.append(len(s_))
            pos += len(s_)
        blob = torch.from_numpy(np.frombuffer(b"".join(strings), dtype=np.uint8).copy()).to(dev)
        t_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        t_len = torch.tensor(lens, dtype=torch.int64, device=dev)
        cdf_d = cdf.detach().to(dev, torch.int32).contiguous()
        len_d = cdf_length.detach().to(dev, torch.int32).contiguous()
        off_d = offset.detach().to(dev, torch.int32).contiguous()
        t_hat = torch.empty((B, h, w, self.C), dtype=torch.float32, device=dev)
        idx = torch.empty((B, h, w, self.C), dtype=torch.int32, device=dev)
        params = torch.empty((B, h, w, 2 * self.C), dtype=torch.float32, device=dev)
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        ws = self._workspace(d)
        _lib.check(self.lib.stemb200_ar_decode(
            C.byref(d), self.packed.data_ptr(), e0.data_ptr(), table.data_ptr(), blob.data_ptr(), t_off.data_ptr(),
            t_len.data_ptr(), cdf_d.data_ptr(), cdf_d.shape[0], cdf_d.shape[1], len_d.data_ptr(), off_d.data_ptr(),
            int(cdf_length.detach().sum().item()), t_hat.data_ptr(), None, idx.data_ptr(), params.data_ptr(),
            status.data_ptr(),
            ws.data_ptr(), _stream()), "ar_decode")
        self._check_completed(ws, "ar_decode")
        if int(status.max().item()) != 0:
            raise _lib.StemLibError("ar_decode: corrupt rANS stream")
        return t_hat, params


# ------------------------------------------------------------------------------------------------------
# g_a / g_s of the I-frame model (mbt2018), the transforms a P-frame goes through
# ------------------------------------------------------------------------------------------------------
def _gdn_fold(beta_p: Tensor, gamma_p: Tensor) -> Tuple[Tensor, Tensor]:
    """NonNegativeParametrizer forward folded once at load (ops/parametrizers.py:42-45, gdn.py:55-57)."""
    ped = 2.0 ** -36
    beta = torch.clamp_min(beta_p.float(), (1e-6 + ped) ** 0.5) ** 2 - ped
    gamma = torch.clamp_min(gamma_p.float(), ped ** 0.5) ** 2 - ped
    return beta, gamma


class TransformsEngine:
    """Packed g_a / g_s (N = M = 192 for mbt2018 q4; any multiple of 64 works)."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device):
        dev = device
        g = lambda k: sd[k].detach().to(dev, torch.float32)
        self.device = dev
        self.ws = Workspace(dev)
        N = sd["g_a.0.weight"].shape[0]
        M = sd["g_a.6.weight"].shape[0]
        self.N, self.M = N, M
        # --- analysis
        # first layer (3 -> N, k5 s2): row_taps mode, one K step per kernel row straight from a zero-bordered NHWC8
        # canvas of the frame (no im2col buffer); the weight is zero-padded to 8 input channels
        w0 = F.pad(g("g_a.0.weight"), (0, 0, 0, 0, 0, 5)).contiguous()  # (N, 8, 5, 5)
        if N != 192:
            raise ValueError("the fused conv+GDN kernel is built for N = 192 channels (mbt2018 quality 1-8)")

        def gdn_of(name, inverse):
            beta, gamma = _gdn_fold(g(f"{name}.beta"), g(f"{name}.gamma"))
            return (beta, gamma, inverse)

        # STEMB200_FIRST=rowtaps selects the round-1 kernel (NHWC8 canvas, weights streamed per tile) for A/B
        self.first_resident = os.environ.get("STEMB200_FIRST", "resident") != "rowtaps"
        if self.first_resident:
            beta0, gamma0, _ = gdn_of("g_a.1", False)
            self.ga_conv = [FirstLayerOp(g("g_a.0.weight"), g("g_a.0.bias"), beta0, gamma0)]
        else:
            self.ga_conv = [ConvOp(w0, g("g_a.0.bias"), c_in=[8], c_out=N, k=5, stride=2, row_taps=True,
                                   gdn=gdn_of("g_a.1", False),
                                   alg_flops_per_out_pixel=2.0 * 75 * N + 2.0 * N * N)]
        for i in (2, 4):
            self.ga_conv.append(ConvOp(g(f"g_a.{i}.weight"), g(f"g_a.{i}.bias"), c_in=[N], c_out=N, k=5, stride=2,
                                       gdn=gdn_of(f"g_a.{i + 1}", False)))
        self.ga_conv.append(ConvOp(g("g_a.6.weight"), g("g_a.6.bias"), c_in=[N], c_out=M, k=5, stride=2,
                                   out_dtype=DT_F32))
        # --- synthesis
        self.gs_conv = []
        for i, cin in ((0, M), (2, N), (4, N)):
            self.gs_conv.append(ConvOp(g(f"g_s.{i}.weight"), g(f"g_s.{i}.bias"), c_in=[cin], c_out=N, k=5, stride=2,
                                       transposed=True, gdn=gdn_of(f"g_s.{i + 1}", True)))
        # last deconv (N -> 3, k5 s2 p2 op1) as a stride-2 conv over 2x2 input super pixels: output super pixel
        # (i, j) holds the 4x4x3 block x_hat[c][4i+u][4j+v]; it reads input rows 2i + a, a = R - 2 in {-1, 0, 1, 2}
        # (k5 taps with R = 0 / S = 0 masked), and the deconv tap is r = u + 2 - 2a = u + 6 - 2R (0 outside 0..4).
        # Each input tile is then fetched for 16 taps per 4 pixels instead of 9 taps per pixel (L2 traffic / 2.25).
        wt = g("g_s.6.weight")  # (N, 3, 5, 5) ConvTranspose layout (in, out, kh, kw)
        wm = torch.zeros((64, N, 5, 5), device=dev)
        bm = torch.zeros(64, device=dev)
        b6 = g("g_s.6.bias")
        mask = 0
        for R in range(1, 5):
            for S in range(1, 5):
                mask |= 1 << (R * 5 + S)
                for u in range(4):
                    r = u + 6 - 2 * R
                    if r < 0 or r > 4:
                        continue
                    for v in range(4):
                        s_ = v + 6 - 2 * S
                        if s_ < 0 or s_ > 4:
                            continue
                        o = (u * 4 + v) * 3
                        wm[o:o + 3, :, R, S] = wt[:, :, r, s_].t()
        for uv in range(16):
            bm[uv * 3:uv * 3 + 3] = b6
        self.gs_last = ConvOp(wm, bm, c_in=[N], c_out=64, k=5, stride=2, tap_mask=mask, out_dtype=DT_F32,
                              alg_flops_per_out_pixel=4 * 2.0 * N * 3 * 25)  # 4 input pixels per super pixel
        # default path: the same layer as a per-pixel GEMM fused behind gs4's IGDN epilogue,
        # W6[col_index(r, s, c)][ci] = w[ci][c][r][s] (75 of 96 rows, quad-grouped order of the col2im kernel),
        # followed by that col2im kernel (STEMB200_FUSE_LAST=0 selects the stand-alone merged-phase conv above)
        self.fuse_last = os.environ.get("STEMB200_FUSE_LAST", "1") != "0" and N == 192
        lib = _lib.load()
        w6 = torch.zeros((96, N), device=dev)
        for r in range(5):
            for s_ in range(5):
                for c in range(3):
                    k = lib.stemb200_synthesis_col_index(r, s_, c)
                    if k < 0:
                        _lib.check(k, "synthesis_col_index")
                    w6[k] = wt[:, c, r, s_]
        w6 = w6.reshape(96, N, 1, 1).contiguous()
        self.w6 = ConvOp(w6, torch.zeros(96, device=dev), c_in=[N], c_out=96, k=1, direct_store=True).packed  # pack only
        self.b6 = b6.contiguous()
        self.gs_conv[2].alg_flops_per_out_pixel_last = 2.0 * N * 3 * 25  # final deconv, per gs4 output pixel

    # -------------------------------------------------------------------------------------------------
    def analysis(self, x: Tensor, pad: Tuple[int, int, int, int] = (0, 0, 0, 0)) -> Tuple[Tensor, int, int]:
        """x: (B, 3, H, W) NCHW, fp32 or uint8 (8-bit samples, used as v / 255 like torchvision's ToTensor);
        pad = (left, right, top, bottom) zero canvas (evalSTEM.py:96-109).
        Returns y as NHWC fp32 (B, h, w, M) plus (h, w)."""
        _require_cuda(x)
        if x.dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"analysis: frames must be float32 or uint8, got {x.dtype}")
        x = x.contiguous()
        B, _, H, W = x.shape
        left, right, top, bottom = pad
        Hp, Wp = H + top + bottom, W + left + right
        lib, ws, N = _lib.load(), self.ws, self.N
        if Hp % 2 or Wp % 2:
            raise ValueError("analysis needs an even padded frame size")
        border = 2
        cp = 4 if self.first_resident else 8   # channels per canvas pixel
        canvas = ws.get("ga_canvas", (B * (Hp + 2 * border) * (Wp + 2 * border) * cp + 64,), torch.float16)
        if self.first_resident:
            stage = lib.stemb200_frame_u8_to_nhwc4 if x.dtype == torch.uint8 else lib.stemb200_frame_to_nhwc4
        else:
            stage = lib.stemb200_frame_u8_to_nhwc8 if x.dtype == torch.uint8 else lib.stemb200_frame_to_nhwc8
        _lib.check(stage(x.data_ptr(), canvas.data_ptr(), B, 3, H, W, Hp, Wp, top, left, border, _stream()),
                   "frame_to_nhwc8")
        cur, h, w = canvas, Hp, Wp
        for li in range(3):
            conv = self.ga_conv[li]
            ho, wo = conv.out_hw(h, w)
            gb = conv([cur], B, h, w, ws.get(f"ga_g{li}", (B, ho, wo, N), torch.float16))
            cur, h, w = gb, ho, wo
        conv = self.ga_conv[3]
        ho, wo = conv.out_hw(h, w)
        y = ws.get("ga_y", (B, ho, wo, self.M), torch.float32)
        conv([cur], B, h, w, y)
        return y, ho, wo

    def synthesis(self, y_hat16: Tensor, x_ref: Optional[Tensor] = None,
                  pad: Tuple[int, int, int, int] = (0, 0, 0, 0), sq_err: Optional[Tensor] = None,
                  out: Optional[Tensor] = None, clamp: bool = True) -> Tensor:
        """y_hat16: (B, h, w, M) fp16 NHWC -> x_hat (B, 3, 16h, 16w) fp32 NCHW clamped to [0, 1]; when x_ref
        (unpadded frames) is given, sq_err[b] += sum((x_ref - crop(x_hat))^2)."""
        B, h, w, _ = y_hat16.shape
        lib, ws, N = _lib.load(), self.ws, self.N
        cur = y_hat16
        left, right, top, bottom = pad
        href = wref = 0
        ref_u8 = False
        if x_ref is not None:
            if x_ref.dtype not in (torch.float32, torch.uint8):
                raise TypeError(f"synthesis: x_ref must be float32 or uint8, got {x_ref.dtype}")
            href, wref = x_ref.shape[2], x_ref.shape[3]
            ref_u8 = x_ref.dtype == torch.uint8
        for li in range(3):
            ho, wo = 2 * h, 2 * w
            if li == 2 and self.fuse_last:
                col = ws.get("gs_col", (B, ho, wo, 96), torch.float16)
                self.gs_conv[2].call_last([cur], B, h, w, self.w6, col)
                if out is None:
                    out = ws.get("gs_xhat", (B, 3, 2 * ho, 2 * wo), torch.float32)
                c2i = lib.stemb200_synthesis_col2im_u8 if ref_u8 else lib.stemb200_synthesis_col2im
                _lib.check(c2i(col.data_ptr(), self.b6.data_ptr(), out.data_ptr(), B, ho, wo, _ptr(x_ref), href, wref, top,
                               left, _ptr(sq_err), int(clamp), _stream()), "synthesis_col2im")
                return out
            gb = self.gs_conv[li]([cur], B, h, w, ws.get(f"gs_g{li}", (B, ho, wo, N), torch.float16))
            cur, h, w = gb, ho, wo
        if h % 2 or w % 2:
            raise ValueError("synthesis needs an even latent size at the last layer")
        merged = ws.get("gs_merged", (B, h // 2, w // 2, 64), torch.float32)
        self.gs_last([cur], B, h, w, merged)
        if out is None:
            out = ws.get("gs_xhat", (B, 3, 2 * h, 2 * w), torch.float32)
        tail = lib.stemb200_synthesis_tail_u8 if ref_u8 else lib.stemb200_synthesis_tail
        _lib.check(tail(merged.data_ptr(), out.data_ptr(), B, h // 2, w // 2, _ptr(x_ref), href, wref, top, left,
                        _ptr(sq_err), int(clamp), _stream()), "synthesis_tail")
        return out


# ------------------------------------------------------------------------------------------------------
# STEM entropy model
# ------------------------------------------------------------------------------------------------------
class StemEngine:
    """Packed TPM / HE / HD / context_prediction / EPM + EntropyBottleneck parameters of one STEM variant."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device, has_tpm: bool, has_spm: bool, residual: bool,
                 eb_packed: Tensor, scale_table: Optional[Tensor], scale_bound: float = 0.11,
                 lik_bound: float = 1e-9):
        dev = device
        g = lambda k: sd[k].detach().to(dev, torch.float32)
        self.device = dev
        self.ws = Workspace(dev)
        self.has_tpm, self.has_spm, self.residual = has_tpm, has_spm, residual
        self.scale_bound, self.lik_bound = float(scale_bound), float(lik_bound)
        C2 = sd["HD.4.weight"].shape[0]
        self.C = C2 // 2
        Cc = self.C
        self.zc = sd["HE.4.weight"].shape[0]
        lre = 0.01  # nn.LeakyReLU() default
        if has_tpm:
            self.tpm = [
                ConvOp(g("TPM.0.weight"), g("TPM.0.bias"), c_in=[Cc], c_out=256, k=5, slope=lre),
                ConvOp(g("TPM.2.weight"), g("TPM.2.bias"), c_in=[256], c_out=320, k=5, slope=lre),
                ConvOp(g("TPM.4.weight"), g("TPM.4.bias"), c_in=[320], c_out=C2, k=5),
            ]
        self.he = [
            ConvOp(g("HE.0.weight"), g("HE.0.bias"), c_in=[Cc, Cc], c_out=256, k=3, slope=lre),
            ConvOp(g("HE.2.weight"), g("HE.2.bias"), c_in=[256], c_out=256, k=5, stride=2, slope=lre),
            ConvOp(g("HE.4.weight"), g("HE.4.bias"), c_in=[256], c_out=self.zc, k=5, stride=2, out_dtype=DT_F32),
        ]
        self.hd = [
            ConvOp(g("HD.0.weight"), g("HD.0.bias"), c_in=[self.zc], c_out=256, k=5, stride=2, transposed=True,
                   slope=lre),
            ConvOp(g("HD.2.weight"), g("HD.2.bias"), c_in=[256], c_out=256, k=5, stride=2, transposed=True,
                   slope=lre),
            ConvOp(g("HD.4.weight"), g("HD.4.bias"), c_in=[256], c_out=C2, k=3),
        ]
        if has_spm:
            # MaskedConv2d: weight * mask, only the 12 causal taps are packed (layers.py:38-47)
            self.ctx = ConvOp(g("context_prediction.weight"), g("context_prediction.bias"), c_in=[Cc], c_out=C2, k=5,
                              tap_mask=MASK_A_5x5)
        n_in = [C2] * (1 + int(has_tpm) + int(has_spm))  # cat order: tp | hp | ctx (:576), hp | ctx (:301), tp | hp (:185)
        self.epm = [
            ConvOp(g("EPM.0.weight"), g("EPM.0.bias"), c_in=n_in, c_out=768, k=1, slope=lre),
            ConvOp(g("EPM.2.weight"), g("EPM.2.bias"), c_in=[768], c_out=576, k=1, slope=lre),
            ConvOp(g("EPM.4.weight"), g("EPM.4.bias"), c_in=[576], c_out=C2, k=1, out_dtype=DT_F32),
        ]
        # EPM.4 with its rows interleaved per 64 channels (sigma 64 | mu 64) for the fused EPM.4 + GaussianConditional kernel
        w4, b4 = g("EPM.4.weight"), g("EPM.4.bias")
        order = torch.cat([torch.cat([torch.arange(c0, c0 + 64), torch.arange(Cc + c0, Cc + c0 + 64)])
                           for c0 in range(0, Cc, 64)]).to(dev) if Cc % 64 == 0 else None
        self.epm4_gc = None if order is None else ConvOp(w4[order].contiguous(), b4[order].contiguous(), c_in=[576],
                                                         c_out=C2, k=1, out_dtype=DT_F32)
        self._ar = None
        if has_spm:
            self._ar_spec = (g("context_prediction.weight") * g("context_prediction.mask"), g("context_prediction.bias"),
                             [(g(f"EPM.{i}.weight"), g(f"EPM.{i}.bias")) for i in (0, 2, 4)],
                             [C2] * (1 + int(has_tpm)))
        self.eb_params = eb_packed.detach().to(dev, torch.float32).contiguous()
        self.scale_table = None if scale_table is None or scale_table.numel() == 0 else \
            scale_table.detach().to(dev, torch.float32).contiguous()
        self._side = None

    # -------------------------------------------------------------------------------------------------
    def hyper_latent(self, y16: Tensor, cond16: Tensor, B: int, h: int, w: int) -> Tensor:
        """z = HE(cat[y_cur, y_cond]) as NHWC fp32 (B, h/4, w/4, zc)  (spatiotemporalpriors.py:562)."""
        ws = self.ws
        f16, f32 = torch.float16, torch.float32
        if h % 4 or w % 4:
            raise ValueError("latent height/width must be multiples of 4 (two stride-2 stages in HE/HD)")
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        srcs = [y16, cond16] if len(self.he[0].c_in) == 2 else [y16]
        t1 = self.he[0](srcs, B, h, w, ws.get("he1", (B, h, w, self.he[0].c_out), f16))
        t2 = self.he[1]([t1], B, h, w, ws.get("he2", (B, h2, w2, self.he[1].c_out), f16))
        return self.he[2]([t2], B, h2, w2, ws.get("z", (B, h4, w4, self.zc), f32))

    def temporal_prior(self, cond16: Tensor, B: int, h: int, w: int) -> Tensor:
        """TPM(y_cond) (spatiotemporalpriors.py:565), NHWC fp16."""
        ws, f16 = self.ws, torch.float16
        p1 = self.tpm[0]([cond16], B, h, w, ws.get("tp1", (B, h, w, 256), f16))
        p2 = self.tpm[1]([p1], B, h, w, ws.get("tp2", (B, h, w, 320), f16))
        return self.tpm[2]([p2], B, h, w, ws.get("tp", (B, h, w, 2 * self.C), f16))

    def hyper_prior(self, zhat16: Tensor, B: int, h: int, w: int) -> Tensor:
        """HD(z_hat) (:564), NHWC fp16."""
        ws, f16 = self.ws, torch.float16
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        c1, c2 = self.hd[0].c_out, self.hd[1].c_out
        d1 = self.hd[0]([zhat16], B, h4, w4, ws.get("hd1", (B, h2, w2, c1), f16))
        d2 = self.hd[1]([d1], B, h2, w2, ws.get("hd2", (B, h, w, c2), f16))
        return self.hd[2]([d2], B, h, w, ws.get("hp", (B, h, w, 2 * self.C), f16))

    def static_priors(self, zhat16: Tensor, cond16: Optional[Tensor], B: int, h: int, w: int) -> List[Tensor]:
        """[TPM(y_cond)] + [HD(z_hat)]: the EPM inputs that do not depend on the frame's own y_hat, NHWC fp16."""
        hp = self.hyper_prior(zhat16, B, h, w)
        return ([self.temporal_prior(cond16, B, h, w)] if self.has_tpm else []) + [hp]

    def ar_head(self) -> "ArHead":
        if not self.has_spm:
            raise RuntimeError("this variant has no spatial context model")
        if self._ar is None:
            self._ar = ArHead(*self._ar_spec, device=self.device, scale_bound=self.scale_bound)
        return self._ar

    def params_from_zhat(self, zhat16: Tensor, cond16: Tensor, yq16: Optional[Tensor], B: int, h: int,
                         w: int) -> Tensor:
        """HD(z_hat), TPM(y_cond), context(y_q), EPM -> (scales | means) NHWC fp32 (B, h, w, 2C)  (:564-577)."""
        ws = self.ws
        f16, f32 = torch.float16, torch.float32
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        srcs = self.static_priors(zhat16, cond16, B, h, w)
        if self.has_spm:
            srcs.append(self.ctx([yq16], B, h, w, ws.get("ctx", (B, h, w, 2 * self.C), f16)))
        return self.entropy_parameters(srcs, B, h, w)

    def gaussian_params(self, y16: Tensor, cond16: Tensor, yq16: Optional[Tensor], B: int, h: int, w: int,
                        z_hat_nchw: Optional[Tensor] = None, z_lik_nchw: Optional[Tensor] = None,
                        bits_z: Optional[Tensor] = None, fused_gc: Optional[dict] = None) -> Optional[Tensor]:
        """HE -> EntropyBottleneck -> HD, TPM, context, EPM. Returns params NHWC fp32 (B, h, w, 2C).
        TPM(y_cond) and context(y_q) depend only on the inputs, HE -> EB -> HD is a chain of small launches (55-148
        tiles at 1080p): the two branches run on two streams and meet at EPM.0 (STEMB200_OVERLAP >= 1)."""
        lib, ws = _lib.load(), self.ws
        f16 = torch.float16
        side = None
        tp = ctx = None
        if overlap_level() >= 1 and (self.has_tpm or self.has_spm):
            if self._side is None:
                self._side = SideStream(self.device)
            side = self._side
            with side:
                if self.has_tpm:
                    tp = self.temporal_prior(cond16, B, h, w)
                if self.has_spm:
                    ctx = self.ctx([yq16], B, h, w, ws.get("ctx", (B, h, w, 2 * self.C), f16))
        z = self.hyper_latent(y16, cond16, B, h, w)
        h4, w4 = h // 4, w // 4
        zhat16 = ws.get("zhat16", (B, h4, w4, self.zc), f16)
        _lib.check(lib.stemb200_entropy_bottleneck_fwd(z.data_ptr(), self.eb_params.data_ptr(), B, self.zc, h4, w4,
                                                       self.lik_bound, zhat16.data_ptr(), _ptr(z_hat_nchw),
                                                       _ptr(z_lik_nchw), _ptr(bits_z), _stream()),
                   "entropy_bottleneck_fwd")
        if side is None:
            srcs = self.static_priors(zhat16, cond16, B, h, w)
            if self.has_spm:
                srcs.append(self.ctx([yq16], B, h, w, ws.get("ctx", (B, h, w, 2 * self.C), f16)))
            return self.entropy_parameters(srcs, B, h, w, fused_gc)
        hp = self.hyper_prior(zhat16, B, h, w)
        side.join()
        return self.entropy_parameters([t for t in (tp, hp, ctx) if t is not None], B, h, w, fused_gc)

    def entropy_parameters(self, srcs: Sequence[Tensor], B: int, h: int, w: int,
                           fused_gc: Optional[dict] = None) -> Optional[Tensor]:
        """EPM on cat(tp | hp | ctx) (:576-577) -> (scales | means) NHWC fp32.  fused_gc = {"y": NHWC fp32, "cond16",
        "y_hat", "lik", "bits"}: the last layer runs with GaussianConditional in its epilogue (:578-579,
        stemb200_conv2d_gc_fwd), nothing is returned and sigma / mu never exist in HBM."""
        ws, f16 = self.ws, torch.float16
        e1 = self.epm[0](list(srcs), B, h, w, ws.get("e1", (B, h, w, self.epm[0].c_out), f16))
        e2 = self.epm[1]([e1], B, h, w, ws.get("e2", (B, h, w, self.epm[1].c_out), f16))
        if fused_gc is None:
            return self.epm[2]([e2], B, h, w, ws.get("gparams", (B, h, w, 2 * self.C), torch.float32))
        op = self.epm4_gc
        if op is None:
            raise ValueError("the fused EPM.4 + GaussianConditional kernel needs a channel count that is a multiple of 64")
        d = op.desc
        d.batch, d.h_in, d.w_in = B, h, w
        arr = (C.c_void_p * 1)(e2.data_ptr())
        _lib.check(op.lib.stemb200_conv2d_gc_fwd(
            C.byref(d), arr, op.packed.data_ptr(), op.bias.data_ptr(), fused_gc["y"].data_ptr(),
            _ptr(fused_gc.get("cond16") if self.residual else None), self.scale_bound, self.lik_bound,
            1 if self.has_spm else 0, _ptr(fused_gc.get("y_hat")), _ptr(fused_gc.get("lik")), _ptr(fused_gc.get("bits")),
            _stream()), "conv2d_gc_fwd")
        return None

    def gaussian_conditional(self, y: Tensor, y_is_nchw: bool, cond16: Optional[Tensor], params: Tensor, B: int,
                             h: int, w: int, y_hat: Optional[Tensor], lik: Optional[Tensor],
                             idx: Optional[Tensor] = None, sym: Optional[Tensor] = None,
                             bits: Optional[Tensor] = None, cond32_nchw: Optional[Tensor] = None) -> None:
        """cond32_nchw (only with an NCHW y): the _Res conditioning latent as the API hands it over, fp32 NCHW - the
        residual y - cond and y_hat = round(y - cond) + cond are then formed in fp32 like the reference does."""
        if idx is not None and self.scale_table is None:
            raise ValueError("build_indexes needs the scale table: call update() first")
        n_scales = 0 if self.scale_table is None else self.scale_table.numel()
        if self.residual and cond32_nchw is not None:
            if not y_is_nchw:
                raise ValueError("cond32_nchw needs an NCHW y")
            _lib.check(_lib.load().stemb200_gaussian_conditional_fwd_cond32(
                y.data_ptr(), cond32_nchw.data_ptr(), params.data_ptr(), B, self.C, h, w, _ptr(self.scale_table),
                n_scales, self.scale_bound, self.lik_bound, 1 if self.has_spm else 0, _ptr(y_hat), _ptr(lik), _ptr(idx),
                _ptr(sym), _ptr(bits), _stream()), "gaussian_conditional_fwd_cond32")
            return
        _lib.check(_lib.load().stemb200_gaussian_conditional_fwd(
            y.data_ptr(), int(y_is_nchw), _ptr(cond16 if self.residual else None), params.data_ptr(), B, self.C, h, w,
            _ptr(self.scale_table), n_scales, self.scale_bound, self.lik_bound, 1 if self.has_spm else 0,
            _ptr(y_hat), _ptr(lik), _ptr(idx), _ptr(sym), _ptr(bits), _stream()), "gaussian_conditional_fwd")

    # -------------------------------------------------------------------------------------------------
    def forward_nchw(self, y_cur: Tensor, y_cond: Tensor, want_indexes: bool = False):
        """Drop-in forward: NCHW fp32 in, NCHW fp32 out (spatiotemporalpriors.py:561-585 & variants)."""
        _require_cuda(y_cur, y_cond)
        y_cur = y_cur.contiguous().float()
        y_cond = y_cond.contiguous().float()
        B, Cc, h, w = y_cur.shape
        if Cc != self.C or tuple(y_cond.shape) != tuple(y_cur.shape):
            raise ValueError(f"expected two (B, {self.C}, h, w) latents, got {tuple(y_cur.shape)} / "
                             f"{tuple(y_cond.shape)}")
        ws, dev = self.ws, self.device
        f16 = torch.float16
        y16 = nchw_to_nhwc_f16(y_cur, ws.get("y16", (B, h, w, Cc), f16))
        cond16 = nchw_to_nhwc_f16(y_cond, ws.get("cond16", (B, h, w, Cc), f16))
        yq16 = None
        if self.has_spm:
            yq16 = nchw_to_nhwc_f16(y_cur, ws.get("yq16", (B, h, w, Cc), f16),
                                    sub=y_cond if self.residual else None, round_first=True)
        zc, h4, w4 = self.zc, h // 4, w // 4
        z_hat = torch.empty((B, zc, h4, w4), dtype=torch.float32, device=dev)
        z_lik = torch.empty((B, zc, h4, w4), dtype=torch.float32, device=dev)
        bits = torch.zeros((2, B), dtype=torch.float64, device=dev)
        params = self.gaussian_params(y16, cond16, yq16, B, h, w, z_hat, z_lik, bits[1])
        y_hat = torch.empty_like(y_cur)
        y_lik = torch.empty_like(y_cur)
        idx = sym = None
        if want_indexes:
            idx = torch.empty(y_cur.shape, dtype=torch.int32, device=dev)
            sym = torch.empty(y_cur.shape, dtype=torch.int32, device=dev)
        # _Res: one definition of the residual everywhere - fp32 y_cur - y_cond, as yq16 above and as compress() use
        self.gaussian_conditional(y_cur, True, cond16, params, B, h, w, y_hat, y_lik, idx, sym, bits[0],
                                  cond32_nchw=y_cond if self.residual else None)
        return {"y_hat": y_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "z_hat": z_hat, "bits": bits,
                "indexes": idx, "symbols": sym, "params_nhwc": params}


class IFrameEntropyEngine(StemEngine):
    """h_a / h_s / context_prediction / entropy_parameters of JointAutoregressiveHierarchicalPriors
    (priors.py:441-470): the STEM WithoutTPM topology without a conditioning latent and with the mbt2018 widths
    (N, N, N | M, 3M/2, 2M | 4M -> 10M/3 -> 8M/3 -> 2M). h_s.2's 3M/2 = 288 outputs are zero-padded to 320 so that
    they tile the 160-wide MMA N block."""

    def __init__(self, sd: Dict[str, Tensor], device: torch.device, eb_packed: Tensor, scale_table: Optional[Tensor],
                 scale_bound: float = 0.11, lik_bound: float = 1e-9):
        dev = device
        g = lambda k: sd[k].detach().to(dev, torch.float32)
        self.device = dev
        self.ws = Workspace(dev)
        self.has_tpm, self.has_spm, self.residual = False, True, False
        self.scale_bound, self.lik_bound = float(scale_bound), float(lik_bound)
        N = sd["h_a.0.weight"].shape[0]
        M = sd["h_a.0.weight"].shape[1]
        self.C, self.zc = M, N
        C2 = 2 * M
        lre = 0.01
        self.he = [
            ConvOp(g("h_a.0.weight"), g("h_a.0.bias"), c_in=[M], c_out=N, k=3, slope=lre),
            ConvOp(g("h_a.2.weight"), g("h_a.2.bias"), c_in=[N], c_out=N, k=5, stride=2, slope=lre),
            ConvOp(g("h_a.4.weight"), g("h_a.4.bias"), c_in=[N], c_out=N, k=5, stride=2, out_dtype=DT_F32),
        ]
        mid = sd["h_s.2.weight"].shape[1]  # 3M/2
        mid_p = (mid + 63) // 64 * 64 if mid % 64 else mid
        w2 = F.pad(g("h_s.2.weight"), (0, 0, 0, 0, 0, mid_p - mid))      # ConvTranspose: (in, out, kh, kw)
        b2 = F.pad(g("h_s.2.bias"), (0, mid_p - mid))
        w4 = F.pad(g("h_s.4.weight"), (0, 0, 0, 0, 0, mid_p - mid, 0, 0))  # Conv: (out, in, kh, kw)
        self.hd = [
            ConvOp(g("h_s.0.weight"), g("h_s.0.bias"), c_in=[N], c_out=M, k=5, stride=2, transposed=True, slope=lre),
            ConvOp(w2.contiguous(), b2, c_in=[M], c_out=mid_p, k=5, stride=2, transposed=True, slope=lre),
            ConvOp(w4.contiguous(), g("h_s.4.bias"), c_in=[mid_p], c_out=C2, k=3),
        ]
        self.ctx = ConvOp(g("context_prediction.weight"), g("context_prediction.bias"), c_in=[M], c_out=C2, k=5,
                          tap_mask=MASK_A_5x5)
        ep = [(g(f"entropy_parameters.{i}.weight"), g(f"entropy_parameters.{i}.bias")) for i in (0, 2, 4)]
        self.epm = [
            ConvOp(ep[0][0], ep[0][1], c_in=[C2, C2], c_out=ep[0][0].shape[0], k=1, slope=lre),
            ConvOp(ep[1][0], ep[1][1], c_in=[ep[0][0].shape[0]], c_out=ep[1][0].shape[0], k=1, slope=lre),
            ConvOp(ep[2][0], ep[2][1], c_in=[ep[1][0].shape[0]], c_out=C2, k=1, out_dtype=DT_F32),
        ]
        self._ar = None
        self._side = None
        self.epm4_gc = None
        self._ar_spec = (g("context_prediction.weight") * g("context_prediction.mask"), g("context_prediction.bias"),
                         ep, [C2])
        self.eb_params = eb_packed.detach().to(dev, torch.float32).contiguous()
        self.scale_table = None if scale_table is None or scale_table.numel() == 0 else \
            scale_table.detach().to(dev, torch.float32).contiguous()


def pad64(h: int, w: int) -> Tuple[int, int, int, int]:
    """(left, right, top, bottom) of evalSTEM.py:96-109."""
    p = 64
    nh, nw = (h + p - 1) // p * p, (w + p - 1) // p * p
    left = (nw - w) // 2
    top = (nh - h) // 2
    return left, nw - w - left, top, nh - h - top


class PFramePipeline:
    """The evalSTEM P-frame loop body (stem/evalSTEM.py:93-154 without the entropy coder) for a whole GOP:
    pad -> g_a -> STEM forward (+likelihoods) -> g_s -> crop/clamp -> bit and squared-error sums.

    For SpatioTemporalPriorModel / _Res the frames of a GOP are processed as one batch (y_hat[t] depends on
    frame t-1 only through an elementwise scan); the WithoutSPM* variants are serial in t.
    """

    def __init__(self, transforms: TransformsEngine, stem: StemEngine):
        self.tr, self.stem = transforms, stem
        self.ws = Workspace(stem.device)
        self._graphed: Dict[tuple, dict] = {}
        self._side = None

    def run_gop(self, frames: Tensor, y_cond0: Tensor, want_outputs: bool = True):
        """`forward_gop` for a stream of GOPs of one shape (an evaluation loop): the kernel chain is captured once
        per shape as a CUDA graph over two static input slots and replayed. `frames` may live on the host (pinned
        memory makes the copy asynchronous): it is copied into the free slot on a copy stream, so the host-to-device
        copy of call i+1 overlaps the kernels of call i; device-resident frames are copied on the current stream.
        Everything is stream-ordered, nothing synchronises the host. The returned tensors are static: `stats`,
        `y_hat`, `lik_*` belong to the slot and are overwritten by the call after next, `x_hat_padded` is a workspace
        buffer overwritten by the next call. Results are bit-identical to `forward_gop`."""
        if y_cond0.device.type != "cuda":
            raise ValueError("run_gop: y_cond0 must be a CUDA tensor")
        key = (tuple(frames.shape), frames.dtype, tuple(y_cond0.shape), bool(want_outputs))
        st = self._graphed.get(key)
        if st is None:
            st = self._graphed[key] = self._capture(frames, y_cond0, want_outputs)
        slot = st["slots"][st["next"]]
        st["next"] ^= 1
        cur = torch.cuda.current_stream()
        if frames.is_cuda:
            slot["frames"].copy_(frames, non_blocking=True)
        else:
            cs = st["copy_stream"]
            cs.wait_event(slot["done"])  # the replay that last read this slot has finished
            with torch.cuda.stream(cs):
                slot["frames"].copy_(frames, non_blocking=True)
                slot["ready"].record(cs)
            cur.wait_event(slot["ready"])
        slot["cond"].copy_(y_cond0, non_blocking=True)
        slot["graph"].replay()
        slot["done"].record(cur)
        return slot["outs"]

    def _capture(self, frames: Tensor, y_cond0: Tensor, want_outputs: bool) -> dict:
        dev = self.stem.device
        slots = []
        for _ in range(2):
            fr = torch.empty(tuple(frames.shape), dtype=frames.dtype, device=dev)
            cd = torch.empty(tuple(y_cond0.shape), dtype=torch.float32, device=dev)
            fr.copy_(frames)
            cd.copy_(y_cond0)
            slots.append({"frames": fr, "cond": cd, "ready": torch.cuda.Event(), "done": torch.cuda.Event()})
        # one eager pass first: workspaces, function attributes and cluster occupancy queries are set up outside capture
        self.forward_gop(slots[0]["frames"], slots[0]["cond"], want_outputs)
        torch.cuda.synchronize()
        for sl in slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                sl["outs"] = self.forward_gop(sl["frames"], sl["cond"], want_outputs)
            sl["graph"] = g
        # the graphs address the workspace buffers of this moment: keep them alive even if a later call with another
        # shape makes the workspaces re-allocate under the same names
        keep = [list(w._bufs.values()) for w in (self.ws, self.tr.ws, self.stem.ws)]
        return {"slots": slots, "next": 0, "copy_stream": torch.cuda.Stream(device=dev), "keep": keep}

    def forward_gop(self, frames: Tensor, y_cond0: Tensor, want_outputs: bool = True):
        """frames: (T, 3, H, W) NCHW CUDA (unpadded), fp32 in [0, 1] or uint8 (8-bit samples, evaluated as v / 255:
        bit-identical to feeding ToTensor's output, evalSTEM.py:185); y_cond0: (1, C, h, w) fp32 NCHW (previous decoded
        latent). Returns dict with per-frame bits_y, bits_z, sq_err (fp64 device tensors) and, optionally,
        x_hat / y_hat / likelihood tensors."""
        _require_cuda(frames, y_cond0)
        st, tr, ws = self.stem, self.tr, self.ws
        lib = _lib.load()
        T, _, H, W = frames.shape
        pad = pad64(H, W)
        y32, h, w = tr.analysis(frames, pad)  # (T, h, w, C) fp32 NHWC
        Cc, dev = st.C, st.device
        f16, f32 = torch.float16, torch.float32
        yhat_all = ws.get("yhat_all", (T + 1, h, w, Cc), f16)
        nchw_to_nhwc_f16(y_cond0.contiguous().float(), yhat_all[0:1])
        y16 = ws.get("y16", (T, h, w, Cc), f16)
        yq16 = ws.get("yq16", (T, h, w, Cc), f16) if st.has_spm else None
        stats = torch.zeros((3, T), dtype=torch.float64, device=dev)  # bits_y, bits_z, sq_err
        zc, h4, w4 = st.zc, h // 4, w // 4
        outs = {}
        if want_outputs:
            outs["y_hat"] = torch.empty((T, Cc, h, w), dtype=f32, device=dev)
            outs["lik_y"] = torch.empty((T, Cc, h, w), dtype=f32, device=dev)
            outs["lik_z"] = torch.empty((T, zc, h4, w4), dtype=f32, device=dev)
        per = h * w * Cc
        if st.has_spm:
            # y_hat scan (elementwise, serial in t only for _Res), then everything else batched over the GOP
            if st.residual:
                for t in range(T):
                    _lib.check(lib.stemb200_latent_stage(y32[t].data_ptr(), yhat_all[t].data_ptr(), y16[t].data_ptr(),
                                                         yq16[t].data_ptr(), yhat_all[t + 1].data_ptr(), per,
                                                         _stream()), "latent_stage")
            else:
                _lib.check(lib.stemb200_latent_stage(y32.data_ptr(), None, y16.data_ptr(), yq16.data_ptr(),
                                                     yhat_all[1:].data_ptr(), per * T, _stream()), "latent_stage")
            cond16 = yhat_all[0:T]
            side = None
            if overlap_level() >= 2:
                # y_hat = round(y [- cond]) [+ cond] is complete: g_s (45 % of the step) does not wait for sigma / mu
                if self._side is None:
                    self._side = SideStream(dev)
                side = self._side
                with side:
                    x_hat = tr.synthesis(yhat_all[1:], x_ref=frames.contiguous(), pad=pad, sq_err=stats[2], out=None)
            if fuse_gc_enabled() and st.epm4_gc is not None:
                st.gaussian_params(y16, cond16, yq16, T, h, w, None, outs.get("lik_z"), stats[1],
                                   fused_gc={"y": y32, "cond16": cond16, "y_hat": outs.get("y_hat"),
                                             "lik": outs.get("lik_y"), "bits": stats[0]})
            else:
                params = st.gaussian_params(y16, cond16, yq16, T, h, w, None, outs.get("lik_z"), stats[1])
                st.gaussian_conditional(y32, False, cond16, params, T, h, w, outs.get("y_hat"), outs.get("lik_y"),
                                        bits=stats[0])
            if side is not None:
                side.join()
                outs.update(x_hat_padded=x_hat, pad=pad, stats=stats, num_pixels=H * W)
                return outs
        else:
            # y_hat[t] = round(y - mu) + mu depends on the whole network applied to y_hat[t-1]: serial
            _lib.check(lib.stemb200_latent_stage(y32.data_ptr(), None, y16.data_ptr(), None, None, per * T,
                                                 _stream()), "latent_stage")
            yh = outs.get("y_hat")
            if yh is None:
                yh = ws.get("yhat_nchw", (T, Cc, h, w), f32)
            fuse = fuse_gc_enabled() and st.epm4_gc is not None
            for t in range(T):
                lz = outs["lik_z"][t:t + 1] if want_outputs else None
                ly = outs["lik_y"][t:t + 1] if want_outputs else None
                if fuse:
                    st.gaussian_params(y16[t:t + 1], yhat_all[t:t + 1], None, 1, h, w, None, lz, stats[1, t:t + 1],
                                       fused_gc={"y": y32[t:t + 1], "y_hat": yh[t:t + 1], "lik": ly,
                                                 "bits": stats[0, t:t + 1]})
                else:
                    params = st.gaussian_params(y16[t:t + 1], yhat_all[t:t + 1], None, 1, h, w, None, lz,
                                                stats[1, t:t + 1])
                    st.gaussian_conditional(y32[t:t + 1], False, None, params, 1, h, w, yh[t:t + 1], ly,
                                            bits=stats[0, t:t + 1])
                nchw_to_nhwc_f16(yh[t:t + 1], yhat_all[t + 1:t + 2])
        x_hat = tr.synthesis(yhat_all[1:], x_ref=frames.contiguous(), pad=pad, sq_err=stats[2],
                             out=None)
        outs["x_hat_padded"] = x_hat
        outs["pad"] = pad
        outs["stats"] = stats
        outs["num_pixels"] = H * W
        return outs
