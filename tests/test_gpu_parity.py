"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle and the
reference-generated golden fixtures.

Tolerances (BASELINE.json north_star): symbols / CDF indexes bit-exact given identical scales; likelihoods
within 1e-4 relative given identical inputs; end-to-end bpp within 0.5 %, PSNR within 0.01 dB.  The dense
layers compute with fp16 operands / fp32 accumulation, so per-layer outputs are compared with a relative-RMS
bound (2e-3) rather than element-wise equality.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stem_oracle as O
from spatiotemporalentropymodel_b200 import synthetic as S

pytestmark = pytest.mark.gpu

LIK_RTOL = 1e-4


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_rms(a, b):
    a, b = a.double(), b.double()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


def lik_close(got, ref, sigma=None):
    """Max relative error after the 1e-9 floor ("floor-aware"), over the elements whose scale lies
    inside the reference's scale table (sigma <= 256).  Beyond the table the reference's own fp32 formula
    (difference of two erfc values near 1/2, entropy_models.py:583-585) is only accurate to ~5e-4 relative
    (1 ulp of 0.5 against a likelihood ~ 1/(sigma*sqrt(2*pi))), and torch's CPU erfc is not even the same routine
    for the vectorised body and the scalar tail of one tensor, so there the bound is absolute: 4 ulp(0.5)."""
    got, ref = got.double().flatten(), ref.double().flatten()
    err = (got - ref).abs()
    if sigma is None:
        return float((err / ref.abs()).max())
    inside = sigma.flatten().double() <= 256.0
    assert float(err[~inside].max() if (~inside).any() else 0.0) <= 2.4e-7
    return float((err[inside] / ref[inside].abs()).max())


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from spatiotemporalentropymodel_b200 import _lib
    _lib.load()  # fails loudly if libstemb200.so is missing
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------- a9 isolated
def test_gaussian_conditional_kat_bit_exact(dev, golden):
    from spatiotemporalentropymodel_b200.engine import gaussian_conditional_flat
    g = golden("gaussian_conditional_kat.npz")
    y, mu, sigma = t(g["y"]).to(dev), t(g["mu"]).to(dev), t(g["sigma"]).to(dev)
    table = t(g["scale_table"]).to(dev)
    y_hat, lik, idx, sym, bits = gaussian_conditional_flat(y, sigma, mu, table, 0.11, 1e-9, want_idx=True,
                                                          want_sym=True, want_bits=True)
    assert torch.equal(idx.cpu(), t(g["idx"])), "scale-table indexes must be bit-exact"
    assert torch.equal(sym.cpu(), t(g["sym"])), "symbols must be bit-exact"
    assert torch.equal(y_hat.cpu(), t(g["y_hat"])), "dequantised values must be bit-exact"
    assert lik_close(lik.cpu(), t(g["lik"]), t(g["sigma"])) <= LIK_RTOL
    ref_bits = float((-torch.log2(t(g["lik"]).double())).sum())
    assert abs(float(bits.item()) - ref_bits) / ref_bits < 1e-5


def test_gaussian_conditional_module_api(dev, golden):
    from spatiotemporalentropymodel_b200.entropy_models import GaussianConditional
    from spatiotemporalentropymodel_b200.models import get_scale_table
    g = golden("gaussian_conditional_kat.npz")
    gc = GaussianConditional(None).eval()
    gc.update_scale_table(get_scale_table(), force=True)
    gc = gc.to(dev)
    y, mu, sigma = t(g["y"]).to(dev), t(g["mu"]).to(dev), t(g["sigma"]).to(dev)
    y_hat, lik = gc(y.view(2, 4, 32, 32), sigma.view(2, 4, 32, 32), means=mu.view(2, 4, 32, 32))
    assert y_hat.shape == (2, 4, 32, 32)
    assert torch.equal(y_hat.flatten().cpu(), t(g["y_hat"]))
    assert torch.equal(gc.build_indexes(sigma).cpu(), t(g["idx"]))
    # means=None path: round(x) exactly (compressai_tests/test_entropy_models.py:249-259)
    out, _ = gc(y, sigma)
    assert torch.equal(out.cpu(), torch.round(t(g["y"])))


def test_gaussian_conditional_large_properties(dev):
    """Full 1080p latent size: idempotence of the quantiser, index monotonicity, floor, checksum vs oracle."""
    from spatiotemporalentropymodel_b200.engine import gaussian_conditional_flat
    n = 192 * 68 * 120
    g = torch.Generator().manual_seed(3)
    y = 4 * torch.randn(n, generator=g)
    mu = torch.rand(n, generator=g) * 4 - 2
    sigma = torch.exp(torch.rand(n, generator=g) * 9 - 4)
    table = O.get_scale_table()
    yh, lik, idx, sym, bits = gaussian_conditional_flat(y.to(dev), sigma.to(dev), mu.to(dev), table.to(dev), 0.11,
                                                        1e-9, want_idx=True, want_sym=True, want_bits=True)
    yh2, _, _, sym2, _ = gaussian_conditional_flat(yh, sigma.to(dev), mu.to(dev), None, 0.11, 1e-9, want_sym=True)
    assert torch.equal(yh2, yh) and torch.equal(sym2, sym)          # quantising a quantised value is a no-op
    order = torch.argsort(sigma)
    assert bool((idx.cpu()[order][1:] >= idx.cpu()[order][:-1]).all())  # indexes are monotone in sigma
    assert float(lik.min()) >= 9.99e-10 and float(lik.max()) <= 1.0
    ref_yh, ref_lik = O.gaussian_conditional_forward(y, sigma, mu)
    assert torch.equal(yh.cpu(), ref_yh)
    assert torch.equal(idx.cpu(), O.build_indexes(sigma, table))
    assert int(sym.cpu().long().sum()) == int(O.quantize_symbols(y, mu).long().sum())
    assert lik_close(lik.cpu(), ref_lik, sigma) <= LIK_RTOL
    ref_bits = float((-torch.log2(ref_lik.double())).sum())
    assert abs(float(bits.item()) - ref_bits) / ref_bits < 1e-5


# ------------------------------------------------------------------------------------------- a4
def test_entropy_bottleneck_kat(dev, golden):
    from spatiotemporalentropymodel_b200.entropy_models import EntropyBottleneck
    g = golden("entropy_bottleneck_kat.npz")
    sd = S.make_stem_state_dict("SpatioTemporalPriorModel", seed=0)
    eb = EntropyBottleneck(256)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items()
                        if k.startswith("entropy_bottleneck.") and v.numel()}, strict=False)
    eb = eb.to(dev).eval()
    z_hat, lik = eb(t(g["z"]).to(dev))
    assert torch.equal(z_hat.cpu(), t(g["z_hat"]))
    assert lik_close(lik.cpu(), t(g["lik"])) <= LIK_RTOL


# ------------------------------------------------------------------------------------------- conv layers
CONV_CASES = [
    # name, cin list, cout, k, stride, transposed, masked, h, w
    ("tpm0_5x5", [192], 256, 5, 1, False, False, 16, 24),
    ("he0_3x3_cat", [192, 192], 256, 3, 1, False, False, 16, 24),
    ("he2_5x5_s2", [256], 256, 5, 2, False, False, 16, 24),
    ("hd0_deconv", [256], 256, 5, 2, True, False, 7, 9),
    ("ctx_masked", [192], 384, 5, 1, False, True, 16, 24),
    ("epm0_1x1_cat3", [384, 384, 384], 768, 1, 1, False, False, 16, 24),
    ("tpm2_320", [256], 320, 5, 1, False, False, 12, 20),
    # >= 2 x 148 tiles: pair mode (2-CTA clusters, weight tiles multicast) and the interleaved deconv phases
    ("tpm0_5x5_pairs", [192], 256, 5, 1, False, False, 136, 240),
    ("hd0_deconv_pairs", [256], 256, 5, 2, True, False, 68, 120),
    ("tpm2_320_big", [256], 320, 5, 1, False, False, 136, 240),
    # 3 x 255 tiles: odd tile count, cluster mode with a phantom last tile
    ("tpm0_5x5_phantom", [192], 256, 5, 1, False, False, 136, 240, 3),
    ("hd2_deconv_phantom", [256], 256, 5, 2, True, False, 136, 240, 1),
]


def _run_conv_case(dev, case, x, wt, bias):
    from spatiotemporalentropymodel_b200.engine import ConvOp, MASK_A_5x5, nchw_to_nhwc_f16, nhwc_f32_to_nchw
    from spatiotemporalentropymodel_b200._lib import DT_F32
    name, cins, cout, k, stride, transposed, masked, h, w = case[:9]
    B = x.shape[0]
    if masked:
        ref = F.conv2d(x, O.masked_weight({"context_prediction.weight": wt}), bias, padding=2)
    elif transposed:
        ref = F.conv_transpose2d(x, wt, bias, stride=2, padding=k // 2, output_padding=1)
    else:
        ref = F.conv2d(x, wt, bias, stride=stride, padding=k // 2)
    ref = F.leaky_relu(ref, 0.01)
    op = ConvOp(wt.to(dev), bias.to(dev), c_in=cins, c_out=cout, k=k, stride=stride, transposed=transposed,
                tap_mask=MASK_A_5x5 if masked else 0, slope=0.01, out_dtype=DT_F32)
    srcs, c0 = [], 0
    for c in cins:
        xs = x[:, c0:c0 + c].contiguous().to(dev)
        srcs.append(nchw_to_nhwc_f16(xs, torch.empty((B, h, w, c), dtype=torch.float16, device=dev)))
        c0 += c
    ho, wo = op.out_hw(h, w)
    out = op(srcs, B, h, w, torch.full((B, ho, wo, cout), float("nan"), device=dev))
    got = nhwc_f32_to_nchw(out, torch.empty((B, cout, ho, wo), device=dev)).cpu()
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    return got, ref


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_layers_vs_torch(dev, case):
    """Each geometry of the path through stemb200_conv2d_fwd against F.conv2d / F.conv_transpose2d (fp32, CPU).
    (a) fp16-representable operands: the only difference left is the fp32 accumulation order - tight absolute bound
    (catches any indexing / tap / padding error); (b) arbitrary fp32 operands, as a checkpoint and real activations
    have them: the operands are rounded to fp16 on the way to the tensor core (kind::f16, 11-bit significands), and
    the result must stay within the error that rounding allows - relative RMS ~ 2^-11 * sqrt(2/3) ~ 4e-4 against the
    exact fp32 convolution, with no bias (mean error << RMS error)."""
    name, cins, cout, k, stride, transposed, masked, h, w = case[:9]
    g = torch.Generator().manual_seed(11)
    B, cin = (case[9] if len(case) > 9 else 2), sum(cins)
    x32 = torch.randn((B, cin, h, w), generator=g)
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    wt32 = torch.randn(wshape, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    got, ref = _run_conv_case(dev, case, x32.half().float(), wt32.half().float(), bias)
    assert float((got - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))
    got, ref = _run_conv_case(dev, case, x32, wt32, bias)
    d = (got - ref).double()
    noise = ref - bias.reshape(1, -1, 1, 1)   # the part of the output the rounded operands produce
    rel = float(torch.sqrt((d ** 2).mean() / (noise.double() ** 2).mean()))
    assert rel < 6e-4, rel
    assert abs(float(d.mean())) < 0.1 * float(d.std()) + 1e-7, (float(d.mean()), float(d.std()))


@pytest.mark.parametrize("scale", [1.0, 100.0, 400.0])
def test_activation_range_of_the_fused_gdn_layers(dev, scale):
    """fp16 range check: conv + GDN / deconv + IGDN with activations `scale` times larger than the
    synthetic checkpoints produce. The kernel carries x^2 as (x/8)^2 in fp16, so |x| up to ~2047 stays finite; GDN
    output is bounded by 1/sqrt(gamma_ii) whatever the input, IGDN grows quadratically (and a trained IGDN sees the
    small, normalised values GDN made). Outputs must stay finite and within the fp16-operand error of the oracle."""
    from spatiotemporalentropymodel_b200.engine import ConvOp, _gdn_fold, nchw_to_nhwc_f16, nhwc_f16_to_nchw
    g = torch.Generator().manual_seed(5)
    C, B, h, w = 192, 1, 24, 40
    ped = torch.tensor([2.0 ** -36])
    beta_p = torch.sqrt(torch.max(1.0 + 0.5 * torch.rand(C, generator=g) + ped, ped))
    gamma_p = torch.sqrt(torch.max(0.1 * torch.eye(C) + 0.02 * torch.rand((C, C), generator=g) + ped, ped))
    beta, gamma = _gdn_fold(beta_p.to(dev), gamma_p.to(dev))
    for inverse in (False, True):
        # pre-activation |x| ~ scale (GDN) / ~ scale / 30 (IGDN: output ~ x^2 * sqrt(sum gamma) must fit fp16)
        amp = scale if not inverse else scale / 30.0
        x = torch.randn((B, C, h, w), generator=g)
        wt = amp * torch.randn((C, C, 5, 5), generator=g) / math.sqrt(C * 25 / (4 if inverse else 1))
        bias = 0.3 * amp * torch.randn(C, generator=g)
        if inverse:
            pre = F.conv_transpose2d(x.half().float(), wt.half().float(), bias, stride=2, padding=2, output_padding=1)
        else:
            pre = F.conv2d(x.half().float(), wt.half().float(), bias, stride=2, padding=2)
        assert float(pre.abs().max()) < 2000.0          # inside the range the (x/8)^2 trick covers
        ref = O.gdn(pre, beta_p, gamma_p, inverse)
        op = ConvOp(wt.to(dev), bias.to(dev), c_in=[C], c_out=C, k=5, stride=2, transposed=inverse,
                    gdn=(beta, gamma, inverse))
        x16 = nchw_to_nhwc_f16(x.to(dev), torch.empty((B, h, w, C), dtype=torch.float16, device=dev))
        ho, wo = op.out_hw(h, w)
        out = op([x16], B, h, w, torch.full((B, ho, wo, C), float("nan"), dtype=torch.float16, device=dev))
        got = nhwc_f16_to_nchw(out, torch.empty((B, C, ho, wo), device=dev)).cpu()
        assert torch.isfinite(got).all(), (scale, inverse)
        assert float(ref.abs().max()) < 6.0e4
        assert rel_rms(got, ref) < 1.5e-3, (scale, inverse, rel_rms(got, ref))


GDN_CASES = [("conv5s2_gdn", False, False, 20, 28), ("deconv5_igdn", True, True, 9, 13), ("gemm_gdn", None, False, 12, 20),
             # pair mode / many tiles per CTA (double-buffered accumulators, ping-pong groups wrap around)
             ("conv5s2_gdn_pairs", False, False, 272, 480), ("deconv5_igdn_pairs", True, True, 68, 120),
             ("gemm_gdn_pairs", None, False, 160, 256),
             # odd tile count per sub-problem (3 x 255 / 1 x 255 tiles): cluster mode with a phantom last tile
             ("conv5s2_gdn_phantom", False, False, 272, 480, 3), ("deconv5_igdn_phantom", True, True, 136, 240, 1)]


@pytest.mark.parametrize("case", GDN_CASES, ids=[c[0] for c in GDN_CASES])
def test_fused_conv_gdn_vs_oracle(dev, case):
    """stemb200_conv2d_gdn_fwd (conv/deconv + GDN/IGDN in one kernel) against F.conv2d + the oracle's GDN
    (layers/gdn.py:52-67 with the NonNegativeParametrizer reparametrisation)."""
    from spatiotemporalentropymodel_b200.engine import ConvOp, _gdn_fold, nchw_to_nhwc_f16, nhwc_f16_to_nchw
    name, transposed, inverse, h, w = case[:5]
    g = torch.Generator().manual_seed(21)
    C, B = 192, (case[5] if len(case) > 5 else 2)
    cin = 128 if transposed is None else 192
    k = 1 if transposed is None else 5
    x = torch.randn((B, cin, h, w), generator=g).half().float()
    wshape = (cin, C, k, k) if transposed else (C, cin, k, k)
    wt = (1.5 * torch.randn(wshape, generator=g) / math.sqrt(cin * k * k / (4 if transposed else 1))).half().float()
    bias = 0.3 * torch.randn(C, generator=g)
    ped = torch.tensor([2.0 ** -36])
    beta_p = torch.sqrt(torch.max(1.0 + 0.5 * torch.rand(C, generator=g) + ped, ped))
    gamma_p = torch.sqrt(torch.max(0.1 * torch.eye(C) + 0.02 * torch.rand((C, C), generator=g) + ped, ped))
    if transposed:
        pre = F.conv_transpose2d(x, wt, bias, stride=2, padding=2, output_padding=1)
    elif transposed is None:
        pre = F.conv2d(x, wt, bias)
    else:
        pre = F.conv2d(x, wt, bias, stride=2, padding=2)
    ref = O.gdn(pre, beta_p, gamma_p, inverse)
    beta, gamma = _gdn_fold(beta_p.to(dev), gamma_p.to(dev))
    op = ConvOp(wt.to(dev), bias.to(dev), c_in=[cin], c_out=C, k=k, stride=1 if transposed is None else 2,
                transposed=bool(transposed), gdn=(beta, gamma, inverse))
    x16 = nchw_to_nhwc_f16(x.to(dev), torch.empty((B, h, w, cin), dtype=torch.float16, device=dev))
    ho, wo = op.out_hw(h, w)
    out = op([x16], B, h, w, torch.full((B, ho, wo, C), float("nan"), dtype=torch.float16, device=dev))
    got = nhwc_f16_to_nchw(out, torch.empty((B, C, ho, wo), device=dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel_rms(got, ref) < 1.5e-3
    assert float((got - ref).abs().max()) < 6e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("geom", [(2, 40, 72), (1, 34, 60), (3, 272, 480), (1, 128, 130)],
                         ids=["small", "ragged", "many_tiles", "odd_half"])
def test_first_layer_resident_kernel_vs_oracle(dev, geom, monkeypatch):
    """g_a.0 + GDN with resident weights (csrc/conv_first.cu: NHWC4 canvas, 64-byte-swizzled K = 160 operands, W0 and
    gamma loaded once per CTA) against F.conv2d + the oracle's GDN (priors.py:422-423, gdn.py:52-67), and against the
    row_taps kernel it replaces (same fp16 operands, another K order: equal to fp32 summation noise)."""
    from spatiotemporalentropymodel_b200 import _lib
    from spatiotemporalentropymodel_b200.engine import ConvOp, FirstLayerOp, _gdn_fold, SQ_SCALE, nhwc_f16_to_nchw
    B, H, W = geom
    g = torch.Generator().manual_seed(31)
    x = torch.rand((B, 3, H, W), generator=g)
    wt = 3.0 * (torch.rand((192, 3, 5, 5), generator=g) * 2 - 1) * math.sqrt(3.0 / 75)
    bias = 0.1 * torch.randn(192, generator=g)
    ped = torch.tensor([2.0 ** -36])
    beta_p = torch.sqrt(torch.max(1.0 + 0.5 * torch.rand(192, generator=g) + ped, ped))
    gamma_p = torch.sqrt(torch.max(0.1 * torch.eye(192) + 0.02 * torch.rand((192, 192), generator=g) + ped, ped))
    ref = O.gdn(F.conv2d(x.half().float(), wt.half().float(), bias, stride=2, padding=2), beta_p, gamma_p, False)
    beta, gamma = _gdn_fold(beta_p.to(dev), gamma_p.to(dev))
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    ho, wo = H // 2, W // 2
    outs = {}
    for cp in (4, 8):
        canvas = torch.full((B * (H + 4) * (W + 4) * cp + 64,), float("nan"), dtype=torch.float16, device=dev)
        stage = lib.stemb200_frame_to_nhwc4 if cp == 4 else lib.stemb200_frame_to_nhwc8
        _lib.check(stage(x.to(dev).data_ptr(), canvas.data_ptr(), B, 3, H, W, H, W, 0, 0, 2, st), "canvas")
        canvas[-64:] = 0
        if cp == 4:
            op = FirstLayerOp(wt.to(dev), bias.to(dev), beta, gamma)
        else:
            w8 = F.pad(wt, (0, 0, 0, 0, 0, 5)).contiguous().to(dev)
            op = ConvOp(w8, bias.to(dev), c_in=[8], c_out=192, k=5, stride=2, row_taps=True, gdn=(beta, gamma, False))
        out = op([canvas], B, H, W, torch.full((B, ho, wo, 192), float("nan"), dtype=torch.float16, device=dev))
        outs[cp] = nhwc_f16_to_nchw(out, torch.empty((B, 192, ho, wo), device=dev)).cpu()
        assert torch.isfinite(outs[cp]).all(), cp
    assert rel_rms(outs[4], ref) < 1.5e-3, rel_rms(outs[4], ref)
    assert float((outs[4] - ref).abs().max()) < 6e-3 * max(1.0, float(ref.abs().max()))
    assert rel_rms(outs[4], outs[8]) < 6e-4          # both round their result to fp16


# ------------------------------------------------------------------------------------------- STEM forward
CAL_TAG = {"default": "", "lowrate": "lowrate_"}


def frame_range(calibration):
    return (0.0, 1.0) if calibration == "default" else S.LOWRATE_FRAME_RANGE


@pytest.mark.parametrize("calibration", S.CALIBRATIONS)
@pytest.mark.parametrize("variant", S.STEM_VARIANTS)
def test_stem_forward_vs_reference_golden(dev, golden, variant, calibration):
    from spatiotemporalentropymodel_b200 import models as M
    g = golden(f"stem_{CAL_TAG[calibration]}{variant}.npz")
    model = getattr(M, variant)()
    model.load_state_dict(S.make_stem_state_dict(variant, seed=0, calibration=calibration))
    model.update(force=True)
    model = model.to(dev).eval()
    y_cur, y_cond = t(g["y_cur"]).to(dev), t(g["y_cond"]).to(dev)
    out = model(y_cur, y_cond)
    assert set(out.keys()) == {"y_hat", "likelihoods"} and set(out["likelihoods"].keys()) == {"y", "z"}
    ref_lz, ref_ly = t(g["lik_z"]), t(g["lik_y"])
    assert out["likelihoods"]["z"].shape == ref_lz.shape and out["likelihoods"]["y"].shape == ref_ly.shape
    has_spm = "WithoutSPM" not in variant
    if has_spm:
        assert torch.equal(out["y_hat"].cpu(), t(g["y_hat"]))  # round(y [- cond]) [+ cond] is exact
    bits = lambda l: float((-torch.log2(l.double())).sum())
    got_bits = bits(out["likelihoods"]["y"].cpu()) + bits(out["likelihoods"]["z"].cpu())
    ref_bits = bits(ref_ly) + bits(ref_lz)
    assert abs(got_bits - ref_bits) / ref_bits < 5e-3, (got_bits, ref_bits)   # bpp within 0.5 %
    # z: a flipped rounding of z changes z_hat by 1; report the mismatch rate, require it to stay rare
    z_bits_err = abs(bits(out["likelihoods"]["z"].cpu()) - bits(ref_lz)) / bits(ref_lz)
    assert z_bits_err < 5e-3

    # given identical scales/means the entropy kernel itself is exact: feed the engine's own parameters
    # to the oracle's GaussianConditional
    full = model.forward_with_indexes(y_cur, y_cond)
    params = full["params_nhwc"].permute(0, 3, 1, 2).contiguous().cpu()
    scales, means = params.chunk(2, 1)
    target = (t(g["y_cur"]) - t(g["y_cond"])) if variant.endswith("_Res") else t(g["y_cur"])
    ref_yhat, ref_lik = O.gaussian_conditional_forward(target, scales, means)
    assert lik_close(full["likelihoods"]["y"].cpu(), ref_lik, scales) <= LIK_RTOL
    assert torch.equal(full["indexes"].cpu(), O.build_indexes(scales))
    assert torch.equal(full["symbols"].cpu(), O.quantize_symbols(target, means))
    if not has_spm:
        assert torch.equal(full["y_hat"].cpu(), ref_yhat)


@pytest.mark.parametrize("calibration", S.CALIBRATIONS)
def test_transforms_vs_oracle(dev, golden, calibration):
    from spatiotemporalentropymodel_b200 import models as M
    g = golden(f"stem_{CAL_TAG[calibration]}SpatioTemporalPriorModel.npz")
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(S.make_iframe_state_dict(seed=0, calibration=calibration))
    net = net.to(dev).eval()
    lo, hi = frame_range(calibration)
    frames = S.make_frames(2, 256, 256, seed=1234, lo=lo, hi=hi)
    y, y_noisy = net.getY(frames[1:2].to(dev))
    assert rel_rms(y.cpu(), t(g["y_cur"])) < 2e-3
    assert float((torch.round(y).cpu() != torch.round(t(g["y_cur"]))).float().mean()) < 0.02  # rounding flips stay rare
    d = (y_noisy - y).cpu()                                  # quantize(y, "noise"): y + U(-1/2, 1/2) (priors.py:691)
    assert float(d.abs().max()) <= 0.5 and abs(float(d.mean())) < 0.01 and abs(float(d.std()) - 12 ** -0.5) < 0.01
    x_hat = net.getX(t(g["y_hat"]).to(dev))
    ref = t(g["x_hat"])
    assert x_hat.shape == ref.shape and float(x_hat.min()) >= 0 and float(x_hat.max()) <= 1
    assert rel_rms(x_hat.cpu(), ref) < 2e-3
    if calibration == "lowrate":
        # the reconstruction is a real one here (~27 dB): the PSNR of the frame must survive the fp16 transforms
        mse = float(((frames[1:2] - x_hat.cpu()) ** 2).mean())
        assert abs(-10 * math.log10(mse) - float(g["psnr"])) < 0.01


def _models(variant, calibration, dev):
    from spatiotemporalentropymodel_b200 import models as M
    sd_i = S.make_iframe_state_dict(seed=0, calibration=calibration)
    sd_s = S.make_stem_state_dict(variant, seed=0, calibration=calibration)
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(sd_i)
    stem = getattr(M, variant)()
    stem.load_state_dict(sd_s)
    stem.update(force=True)
    net, stem = net.to(dev).eval(), stem.to(dev).eval()
    return net, stem, M.make_pipeline(net, stem), sd_i, sd_s


def _gop_inputs(T, H, W, seed, calibration, sd_i):
    """T P-frames and the previous decoded latent: round(g_a(frame 0)) by the ORACLE (the I-frame codec stand-in of
    tests/golden/make_golden.py), so that y_cond is what the checkpoint's entropy model expects."""
    lo, hi = frame_range(calibration)
    frames = S.make_frames(T + 1, H, W, seed=seed, lo=lo, hi=hi)
    with torch.no_grad():
        xp, _ = O.pad_to_64(frames[0:1])
        y_cond0 = torch.round(O.g_a(xp, sd_i))
    return frames[1:].contiguous(), y_cond0


@pytest.mark.parametrize("calibration", S.CALIBRATIONS)
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res",
                                     "SpatioTemporalPriorModelWithoutSPM"])
def test_pframe_pipeline_bpp_psnr(dev, variant, calibration):
    """End to end (pad -> g_a -> STEM -> g_s -> crop): bpp within 0.5 %, PSNR within 0.01 dB of the oracle on EVERY
    frame of a GOP of 3 P-frames of a non-multiple-of-64 size (exercises the evalSTEM padding), both checkpoints."""
    from oracle import parity as P
    net, stem, pipe, sd_i, sd_s = _models(variant, calibration, dev)
    H, W, T = 120, 200, 3
    frames, y_cond0 = _gop_inputs(T, H, W, 77, calibration, sd_i)
    out = pipe.forward_gop(frames.to(dev), y_cond0.to(dev))
    with torch.no_grad():
        ref = O.gop_forward(frames, y_cond0, sd_i, sd_s, variant, return_params=True)
    rep = P.gop_parity(out, ref, H, W)
    assert rep["ok"], rep
    l, r, tp, b = out["pad"]
    x_hat = out["x_hat_padded"][:, :, tp:tp + H, l:l + W].cpu()
    assert rel_rms(x_hat, torch.cat([o["x_hat"] for o in ref])) < 5e-2
    assert out["y_hat"].shape == (T, 192, 8, 16)


PARITY_1080P = [(v, c) for c in S.CALIBRATIONS for v in ("SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res",
                                                        "SpatioTemporalPriorModelWithoutSPM")]


@pytest.mark.parametrize("variant,calibration", PARITY_1080P, ids=[f"{v[24:] or 'full'}-{c}" for v, c in PARITY_1080P])
def test_parity_at_benchmark_size_1080p(dev, variant, calibration, monkeypatch):
    """The benchmarked workloads at their real size (1080 x 1920, BASELINE.json configs[1-3]) against the oracle's
    evalSTEM loop (stem/evalSTEM.py:93-154, spatiotemporalpriors.py:561-585): per-frame bpp within 0.5 %, PSNR within
    0.01 dB, for both synthetic checkpoints; the serial WithoutSPM chain over 3 frames (y_hat feeds the next frame
    through the whole network). sigma / mu are also compared directly on the lowrate checkpoint, where < 0.1 % of the
    likelihoods are floored and every one of them reacts to an error."""
    from oracle import parity as P
    net, stem, pipe, sd_i, sd_s = _models(variant, calibration, dev)
    H, W = 1080, 1920
    serial = "WithoutSPM" in variant
    T = 3 if serial else 2
    frames, y_cond0 = _gop_inputs(T, H, W, 1234, calibration, sd_i)
    out = pipe.forward_gop(frames.to(dev), y_cond0.to(dev))       # the shipped path (fused EPM.4 + GaussianConditional)
    torch.cuda.synchronize()
    out = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}
    params = None
    if not serial:  # sigma | mu only exist in HBM when the two kernels run separately
        monkeypatch.setenv("STEMB200_FUSE_GC", "0")
        pipe.forward_gop(frames.to(dev), y_cond0.to(dev))
        torch.cuda.synchronize()
        params = pipe.stem.ws._bufs["gparams"]
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    with torch.no_grad():
        ref = O.gop_forward(frames, y_cond0, sd_i, sd_s, variant, return_params=True)
    rep = P.gop_parity(out, ref, H, W, params)
    assert rep["ok"], rep
    if calibration == "lowrate":
        assert max(f["ref_floored_lik_frac"] for f in rep["frames"]) < 0.01
    if not serial:
        assert rep["max_y_hat_mismatch_frac"] < 0.02          # rounding flips of y_hat = round(y) stay rare
        assert rep["max_sigma_rel_rms"] < 1e-2, rep
        assert rep["max_mu_err_over_sigma_rms"] < 0.2, rep


@pytest.mark.parametrize("size", [(64, 64, 1), (119, 201, 2), (181, 333, 1)], ids=["one_tile", "odd", "odd2"])
@pytest.mark.parametrize("u8", [False, True], ids=["f32", "u8"])
def test_ragged_frame_sizes(dev, size, u8):
    """Frame sizes that are not multiples of anything (evalSTEM.py:96-109 pads them to 64; odd widths switch the frame
    kernels to their scalar loads, odd left / top offsets exercise the crop of the squared-error sum), a single-tile
    frame, T = 1: the same gates against the oracle, 8-bit and fp32 input."""
    from oracle import parity as P
    variant, calibration = "SpatioTemporalPriorModel", "lowrate"
    net, stem, pipe, sd_i, sd_s = _models(variant, calibration, dev)
    H, W, T = size
    frames, y_cond0 = _gop_inputs(T, H, W, 5, calibration, sd_i)
    if u8:
        f8 = torch.round(frames * 255).to(torch.uint8)
        frames, inp = f8.float().div(255.0), f8
    else:
        inp = frames
    out = pipe.forward_gop(inp.to(dev), y_cond0.to(dev))
    with torch.no_grad():
        ref = O.gop_forward(frames, y_cond0, sd_i, sd_s, variant)
    rep = P.gop_parity(out, ref, H, W)
    # a 64 x 64 frame has 16 latent pixels: one rounding flip is 0.03 % of its bits - the bpp gate still holds,
    # the PSNR gate is the north-star one
    assert rep["ok"], rep


def test_parity_4k_frame(dev):
    """Largest frame of BASELINE.json (3840 x 2160, padded to 3840 x 2176: 32 640 latent pixels, 522 k first-layer
    pixels per frame) through the full P-frame pipeline, 8-bit input, against the oracle: the same gates. Exercises
    the tile / index arithmetic at four times the 1080p extents."""
    from oracle import parity as P
    variant, calibration = "SpatioTemporalPriorModel", "lowrate"
    net, stem, pipe, sd_i, sd_s = _models(variant, calibration, dev)
    H, W = 2160, 3840
    frames, y_cond0 = _gop_inputs(1, H, W, 99, calibration, sd_i)
    f8 = torch.round(frames * 255).to(torch.uint8)
    out = pipe.forward_gop(f8.to(dev), y_cond0.to(dev))
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    with torch.no_grad():
        ref = O.gop_forward(f8.float().div(255.0), y_cond0, sd_i, sd_s, variant)
    rep = P.gop_parity(out, ref, H, W)
    assert rep["ok"], rep
    assert rep["max_y_hat_mismatch_frac"] < 0.02


@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModelWithoutSPM"])
def test_uint8_frames_bit_identical_to_totensor_frames(dev, variant):
    """8-bit frames (v / 255 on the device) against the fp32 frames torchvision's ToTensor makes of them
    (stem/evalSTEM.py:185): statistics, latents, likelihoods and reconstruction bit for bit - through forward_gop and
    through the streaming run_gop with pinned host frames."""
    net, stem, pipe, _, _ = _models(variant, "default", dev)
    H, W, T = 120, 200, 2
    f8 = torch.round(S.make_frames(T, H, W, seed=31) * 255).to(torch.uint8)
    f32 = f8.to(torch.float32).div(255.0)
    cond = S.make_latent(1, 192, 8, 16, seed=5).to(dev)
    keys = ("stats", "y_hat", "lik_y", "lik_z", "x_hat_padded")
    want = {k: v.clone() for k, v in pipe.forward_gop(f32.to(dev), cond).items() if k in keys}
    got = pipe.forward_gop(f8.to(dev), cond)
    for k in keys:
        assert torch.equal(got[k], want[k]), k
    got = pipe.run_gop(f8.pin_memory(), cond)
    for k in keys:
        assert torch.equal(got[k], want[k]), ("run_gop", k)
    with pytest.raises(TypeError):
        pipe.forward_gop(f32.double().to(dev), cond)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [(2, 6, 10, 90, 150, (5, 5, 3, 3)), (1, 34, 60, 540, 950, (5, 5, 2, 2))],
                         ids=["ragged_small", "pairs"])
def test_fused_last_synthesis_layer_matches_standalone_and_oracle(geom):
    """g_s.4 + IGDN with the final deconv(N, 3) fused behind it (stemb200_conv2d_gdn_last_fwd + col2im) against the
    stand-alone merged-phase conv path and against the oracle's g_s (priors.py:431-439), ragged frame incl. MSE; the
    second size has enough tiles for 2-CTA clusters and many col2im tiles per CTA."""
    import os
    from spatiotemporalentropymodel_b200.engine import TransformsEngine, nchw_to_nhwc_f16
    B, h, w, Hr, Wr, pad = geom
    dev = torch.device("cuda:0")
    sd = S.make_iframe_state_dict(0)
    y_hat = torch.round(S.make_latent(B, 192, h, w, seed=8))
    x_ref = S.make_frames(B, Hr, Wr, seed=3)                     # un-padded frame inside the 16h x 16w canvas
    assert Hr + pad[2] + pad[3] == 16 * h and Wr + pad[0] + pad[1] == 16 * w
    outs = {}
    for fuse in ("1", "0"):
        os.environ["STEMB200_FUSE_LAST"] = fuse
        try:
            eng = TransformsEngine({k: v.to(dev) for k, v in sd.items()}, dev)
        finally:
            os.environ.pop("STEMB200_FUSE_LAST", None)
        assert eng.fuse_last == (fuse == "1")
        y16 = nchw_to_nhwc_f16(y_hat.to(dev), torch.empty((B, h, w, 192), dtype=torch.float16, device=dev))
        sq = torch.zeros(B, dtype=torch.float64, device=dev)
        x = eng.synthesis(y16, x_ref=x_ref.to(dev), pad=pad, sq_err=sq,
                          out=torch.empty((B, 3, 16 * h, 16 * w), device=dev))
        outs[fuse] = (x.cpu(), sq.cpu())
    ref = O.g_s(y_hat, sd, clamp=True)
    for fuse, (x, sq) in outs.items():
        assert float((x - ref).abs().max()) < 4e-3, fuse
        crop = x[:, :, pad[2]:pad[2] + Hr, pad[0]:pad[0] + Wr]
        want = ((x_ref - crop).double() ** 2).sum(dim=(1, 2, 3))
        assert torch.allclose(sq, want, rtol=1e-5), fuse
    assert float((outs["1"][0] - outs["0"][0]).abs().max()) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [(2, 3, 37, 53, 3, 5, 48, 64), (1, 3, 40, 64, 0, 0, 40, 64), (1, 4, 16, 24, 4, 6, 24, 36),
                                  (1, 3, 30, 52, 1, 3, 32, 58)])
@pytest.mark.parametrize("u8", [False, True], ids=["f32", "u8"])
def test_frame_to_nhwc8_canvas(dev, geom, u8):
    """Operand canvas of the first analysis layer (priors.py:422 on the evalSTEM.py:96-109 padded frame): vectorised
    (w % 4 == 0) and scalar paths, every left-offset residue, channels 3..7 and the border zero - bit-exact, from fp32
    frames and from 8-bit frames (v / 255)."""
    from spatiotemporalentropymodel_b200 import _lib
    n, c, h, w, top, left, hp, wp = geom
    border = 2
    x = torch.rand(n, c, h, w, generator=torch.Generator().manual_seed(5))
    canvas = torch.full((n, hp + 2 * border, wp + 2 * border, 8), 7.0, dtype=torch.float16, device=dev)
    lib = _lib.load()
    if u8:
        x8 = torch.round(x * 255).to(torch.uint8).to(dev)
        x = x8.float().div(255.0)
        _lib.check(lib.stemb200_frame_u8_to_nhwc8(x8.data_ptr(), canvas.data_ptr(), n, c, h, w, hp, wp, top, left, border,
                                                  torch.cuda.current_stream().cuda_stream), "frame_u8_to_nhwc8")
    else:
        x = x.to(dev)
        _lib.check(lib.stemb200_frame_to_nhwc8(x.data_ptr(), canvas.data_ptr(), n, c, h, w, hp, wp, top, left, border,
                                               torch.cuda.current_stream().cuda_stream), "frame_to_nhwc8")
    want = torch.zeros_like(canvas)
    want[:, border + top:border + top + h, border + left:border + left + w, :c] = x.permute(0, 2, 3, 1).half()
    assert torch.equal(canvas, want)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModelWithoutSPM"])
def test_run_gop_streams_graph_replays_bit_identical_to_forward_gop(dev, variant):
    """PFramePipeline.run_gop (captured CUDA graph per input slot, host frames copied on a side stream) against the
    eager forward_gop on a stream of different GOPs: same statistics and latents, bit for bit, in call order."""
    from spatiotemporalentropymodel_b200 import models as M
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(S.make_iframe_state_dict(seed=0))
    stem = getattr(M, variant)()
    stem.load_state_dict(S.make_stem_state_dict(variant, seed=0))
    net, stem = net.to(dev).eval(), stem.to(dev).eval()
    H, W, T = 120, 200, 2
    gops = [S.make_frames(T, H, W, seed=100 + i).pin_memory() for i in range(5)]
    conds = [S.make_latent(1, 192, 8, 16, seed=200 + i).to(dev) for i in range(5)]
    pipe = M.make_pipeline(net, stem)
    want = []
    for fr, cd in zip(gops, conds):
        o = pipe.forward_gop(fr.to(dev), cd)
        want.append((o["stats"].clone(), o["y_hat"].clone(), o["lik_y"].clone(), o["x_hat_padded"].clone()))
    for i, (fr, cd) in enumerate(zip(gops, conds)):
        src = fr if i % 2 == 0 else fr.to(dev)            # host (copy stream) and device-resident inputs
        o = pipe.run_gop(src, cd)
        got = (o["stats"].clone(), o["y_hat"].clone(), o["lik_y"].clone(), o["x_hat_padded"].clone())
        for a, b in zip(got, want[i]):
            assert torch.equal(a, b), (variant, i)
    assert len(pipe._graphed) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res",
                                     "SpatioTemporalPriorModelWithoutSPM"])
def test_stream_overlap_levels_are_bit_identical(dev, variant, monkeypatch):
    """STEMB200_OVERLAP = 0 (one stream) / 1 (TPM + context beside HE -> EB -> HD) / 2 (+ synthesis beside the entropy
    model): the same kernels on the same buffers, only their stream assignment changes - identical results, eagerly and
    through the captured graph of run_gop."""
    net, stem, pipe, _, _ = _models(variant, "default", dev)
    H, W, T = 120, 200, 3
    frames = S.make_frames(T, H, W, seed=41).to(dev)
    cond = S.make_latent(1, 192, 8, 16, seed=6).to(dev)
    keys = ("stats", "y_hat", "lik_y", "lik_z", "x_hat_padded")
    want = None
    for level in ("0", "1", "2"):
        monkeypatch.setenv("STEMB200_OVERLAP", level)
        got = {k: v.clone() for k, v in pipe.forward_gop(frames, cond).items() if k in keys}
        torch.cuda.synchronize()
        if want is None:
            want = got
        for k in keys:
            assert torch.equal(got[k], want[k]), (level, k)
    monkeypatch.setenv("STEMB200_OVERLAP", "2")
    for _ in range(3):
        got = pipe.run_gop(frames, cond)
        for k in keys:
            assert torch.equal(got[k], want[k]), ("run_gop", k)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res",
                                     "SpatioTemporalPriorModelWithoutSPM"])
@pytest.mark.parametrize("size", [(120, 200, 3), (272, 480, 2)], ids=["ragged", "many_tiles"])
def test_fused_entropy_parameters_gaussian_conditional_is_bit_identical(dev, variant, size, monkeypatch):
    """BASELINE north_star item 2: entropy_parameters' last layer with GaussianConditional in its epilogue
    (stemb200_conv2d_gc_fwd; sigma and mu never reach HBM) against the two separate kernels - the same MMAs in the same
    order and the same gc_math.cuh arithmetic, so y_hat and the likelihoods must be identical bit for bit (the bit count,
    a sum, to 1e-6)."""
    net, stem, pipe, _, _ = _models(variant, "lowrate", dev)
    H, W, T = size
    lo, hi = S.LOWRATE_FRAME_RANGE
    frames = S.make_frames(T + 1, H, W, seed=43, lo=lo, hi=hi).to(dev)
    from spatiotemporalentropymodel_b200.evaluate import pad_to_64
    y0, _ = net.getY(pad_to_64(frames[0:1])[0])
    cond = torch.round(y0)
    keys = ("stats", "y_hat", "lik_y", "lik_z", "x_hat_padded")
    monkeypatch.setenv("STEMB200_FUSE_GC", "0")
    want = {k: v.clone() for k, v in pipe.forward_gop(frames[1:], cond).items() if k in keys}
    monkeypatch.delenv("STEMB200_FUSE_GC")   # the default: fused
    got = pipe.forward_gop(frames[1:], cond)
    torch.cuda.synchronize()
    assert float(want["lik_y"].min()) > 0 and torch.isfinite(want["stats"]).all()
    for k in keys[1:]:
        assert torch.equal(got[k], want[k]), (k, float((got[k] - want[k]).abs().max()))
    # the bit counts are sums of the same fp32 terms grouped differently (per warp here, per 256-thread block there)
    assert torch.allclose(got["stats"], want["stats"], rtol=1e-6, atol=0), (got["stats"], want["stats"])
