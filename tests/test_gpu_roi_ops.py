"""GPU parity tests (-m gpu) of the operator variants the variable-rate `stem_roi` family needs
(compressai/models/stem_roi.py, stem_utils.py): SFT and residual epilogues, 128-channel fused GDN, C_out = 160,
ragged / concatenated K segments, 3x3 strided and transposed convs, pooling and staging kernels."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import stem_oracle as O

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    a, b = a.double(), b.double()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def to_nhwc16(x, dev):
    from spatiotemporalentropymodel_b200.engine import nchw_to_nhwc_f16
    B, C, h, w = x.shape
    return nchw_to_nhwc_f16(x.contiguous().to(dev), torch.empty((B, h, w, C), dtype=torch.float16, device=dev))


def to_nchw(x):
    from spatiotemporalentropymodel_b200.engine import nhwc_f16_to_nchw
    B, h, w, C = x.shape
    return nhwc_f16_to_nchw(x, torch.empty((B, C, h, w), device=x.device)).cpu()


def rnd(g, shape, scale=1.0):
    return (scale * torch.randn(shape, generator=g)).half().float()


@pytest.mark.parametrize("C", [128, 192, 256])
def test_sft_epilogue(dev, C):
    """out = LeakyReLU(x * (1 + conv_g(a)) + conv_b(a), 0.2)  (stem_utils.py:36-43, :55-63)"""
    from spatiotemporalentropymodel_b200.engine import sft_op
    g = torch.Generator().manual_seed(C)
    B, h, w, nh = 2, 12, 20, 128
    x, a = rnd(g, (B, C, h, w)), rnd(g, (B, nh, h, w))
    wg, wb = rnd(g, (C, nh, 3, 3), 1 / math.sqrt(nh * 9)), rnd(g, (C, nh, 3, 3), 1 / math.sqrt(nh * 9))
    bg, bb = 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ref = F.leaky_relu(x * (1 + F.conv2d(a, wg, bg, padding=1)) + F.conv2d(a, wb, bb, padding=1), 0.2)
    op = sft_op(wg.to(dev), bg.to(dev), wb.to(dev), bb.to(dev), c_in=nh, slope=0.2)
    out = op([to_nhwc16(a, dev)], B, h, w, torch.full((B, h, w, C), float("nan"), dtype=torch.float16, device=dev),
             aux=to_nhwc16(x, dev))
    got = to_nchw(out)
    assert torch.isfinite(got).all() and rel_rms(got, ref) < 1.5e-3


def test_residual_epilogue_and_relu(dev):
    from spatiotemporalentropymodel_b200.engine import ConvOp
    from spatiotemporalentropymodel_b200._lib import EPI_ADD
    g = torch.Generator().manual_seed(5)
    B, C, h, w = 2, 192, 10, 14
    x, r = rnd(g, (B, C, h, w)), rnd(g, (B, C, h, w))
    wt, b = rnd(g, (C, C, 3, 3), 1 / math.sqrt(C * 9)), 0.1 * torch.randn(C, generator=g)
    op = ConvOp(wt.to(dev), b.to(dev), c_in=[C], c_out=C, k=3, epilogue=EPI_ADD)
    out = op([to_nhwc16(x, dev)], B, h, w, torch.empty((B, h, w, C), dtype=torch.float16, device=dev),
             aux=to_nhwc16(r, dev))
    assert rel_rms(to_nchw(out), F.conv2d(x, wt, b, padding=1) + r) < 1.5e-3
    relu = ConvOp(wt.to(dev), b.to(dev), c_in=[C], c_out=C, k=3, slope=0.0)   # nn.ReLU (stem_utils.py:31)
    out = relu([to_nhwc16(x, dev)], B, h, w, torch.empty((B, h, w, C), dtype=torch.float16, device=dev))
    assert rel_rms(to_nchw(out), F.relu(F.conv2d(x, wt, b, padding=1))) < 1.5e-3


def test_cout_160_then_ragged_cin_160(dev):
    """qmap_feature_ga1: conv(192,160,3,1) -> LeakyReLU(0.1) -> conv(160,128,3,1)  (stem_roi.py:379-385)"""
    from spatiotemporalentropymodel_b200.engine import ConvOp
    g = torch.Generator().manual_seed(6)
    B, h, w = 1, 20, 28
    x = rnd(g, (B, 192, h, w))
    w1, b1 = rnd(g, (160, 192, 3, 3), 1 / math.sqrt(192 * 9)), 0.1 * torch.randn(160, generator=g)
    w2, b2 = rnd(g, (128, 160, 3, 3), 1 / math.sqrt(160 * 9)), 0.1 * torch.randn(128, generator=g)
    op1 = ConvOp(w1.to(dev), b1.to(dev), c_in=[192], c_out=160, k=3, slope=0.1)
    op2 = ConvOp(w2.to(dev), b2.to(dev), c_in=[160], c_out=128, k=3)
    mid = op1([to_nhwc16(x, dev)], B, h, w, torch.full((B, h, w, 160), float("nan"), dtype=torch.float16, device=dev))
    ref_mid = F.leaky_relu(F.conv2d(x, w1, b1, padding=1), 0.1)
    assert rel_rms(to_nchw(mid), ref_mid) < 1.5e-3
    out = op2([mid], B, h, w, torch.empty((B, h, w, 128), dtype=torch.float16, device=dev))
    assert rel_rms(to_nchw(out), F.conv2d(ref_mid, w2, b2, padding=1)) < 2e-3


def test_three_ragged_sources(dev):
    """qmap_feature_ha1[0]: conv(2*192 + 1, 128, 3, 1) on cat[Qmap(1), y_cur(192), y_cond(192)] (stem_roi.py:565)"""
    from spatiotemporalentropymodel_b200.engine import ConvOp
    from spatiotemporalentropymodel_b200 import _lib
    g = torch.Generator().manual_seed(7)
    B, h, w, f = 2, 8, 12, 4
    qmap = torch.rand((B, 1, h * f, w * f), generator=g)
    y1, y2 = rnd(g, (B, 192, h, w)), rnd(g, (B, 192, h, w))
    wt, b = rnd(g, (128, 385, 3, 3), 1 / math.sqrt(385 * 9)), 0.1 * torch.randn(128, generator=g)
    q8 = torch.empty((B, h, w, 8), dtype=torch.float16, device=dev)
    qmap_d = qmap.to(dev)  # keep the device copy alive while the kernel reads it
    _lib.check(_lib.load().stemb200_qmap_pool(qmap_d.data_ptr(), q8.data_ptr(), B, h, w, f,
                                              torch.cuda.current_stream().cuda_stream), "qmap_pool")
    qp = F.adaptive_avg_pool2d(qmap, (h, w))
    assert torch.allclose(q8[..., 0].float().cpu(), qp[:, 0], atol=1e-3) and float(q8[..., 1:].abs().max()) == 0
    # weight for the padded source: channel 0 real, 1..7 zero
    wpad = torch.cat([wt[:, :1], torch.zeros(128, 7, 3, 3), wt[:, 1:]], 1)
    op = ConvOp(wpad.to(dev), b.to(dev), c_in=[8, 192, 192], c_out=128, k=3, slope=0.1)
    out = op([q8, to_nhwc16(y1, dev), to_nhwc16(y2, dev)], B, h, w,
             torch.empty((B, h, w, 128), dtype=torch.float16, device=dev))
    ref = F.leaky_relu(F.conv2d(torch.cat([qp.half().float(), y1, y2], 1), wt, b, padding=1), 0.1)
    assert rel_rms(to_nchw(out), ref) < 1.5e-3


@pytest.mark.parametrize("inverse", [False, True])
def test_fused_gdn_128(dev, inverse):
    from spatiotemporalentropymodel_b200.engine import ConvOp, _gdn_fold
    g = torch.Generator().manual_seed(8)
    B, C, h, w = 2, 128, 9, 13
    x = rnd(g, (B, C, h, w))
    ped = torch.tensor([2.0 ** -36])
    beta_p = torch.sqrt(torch.max(1.0 + 0.5 * torch.rand(C, generator=g) + ped, ped))
    gamma_p = torch.sqrt(torch.max(0.1 * torch.eye(C) + 0.02 * torch.rand((C, C), generator=g) + ped, ped))
    if inverse:
        wt = rnd(g, (C, C, 5, 5), 3 / math.sqrt(C * 25))
        pre = F.conv_transpose2d(x, wt, None, stride=2, padding=2, output_padding=1)
    else:
        wt = rnd(g, (C, C, 5, 5), 1.5 / math.sqrt(C * 25))
        pre = F.conv2d(x, wt, None, stride=2, padding=2)
    ref = O.gdn(pre, beta_p, gamma_p, inverse)
    beta, gamma = _gdn_fold(beta_p.to(dev), gamma_p.to(dev))
    op = ConvOp(wt.to(dev), torch.zeros(C, device=dev), c_in=[C], c_out=C, k=5, stride=2, transposed=inverse,
                gdn=(beta, gamma, inverse))
    ho, wo = op.out_hw(h, w)
    out = op([to_nhwc16(x, dev)], B, h, w, torch.full((B, ho, wo, C), float("nan"), dtype=torch.float16, device=dev))
    got = to_nchw(out)
    assert got.shape == ref.shape and rel_rms(got, ref) < 1.5e-3


def test_k3_strided_and_transposed(dev):
    """conv(128,128,3) stride 2 (stem_roi.py:387) and deconv(192,128,3) (stem_roi.py:478)"""
    from spatiotemporalentropymodel_b200.engine import ConvOp
    g = torch.Generator().manual_seed(9)
    B, h, w = 2, 14, 18
    x = rnd(g, (B, 128, h, w))
    wt, b = rnd(g, (128, 128, 3, 3), 1 / math.sqrt(128 * 9)), 0.1 * torch.randn(128, generator=g)
    op = ConvOp(wt.to(dev), b.to(dev), c_in=[128], c_out=128, k=3, stride=2, slope=0.1)
    out = op([to_nhwc16(x, dev)], B, h, w, torch.empty((B, 7, 9, 128), dtype=torch.float16, device=dev))
    assert rel_rms(to_nchw(out), F.leaky_relu(F.conv2d(x, wt, b, stride=2, padding=1), 0.1)) < 1.5e-3
    x2 = rnd(g, (B, 192, h, w))
    wt2 = rnd(g, (192, 128, 3, 3), 1 / math.sqrt(192 * 9 / 4))
    op2 = ConvOp(wt2.to(dev), b.to(dev), c_in=[192], c_out=128, k=3, stride=2, transposed=True)
    out2 = op2([to_nhwc16(x2, dev)], B, h, w, torch.empty((B, 28, 36, 128), dtype=torch.float16, device=dev))
    assert rel_rms(to_nchw(out2), F.conv_transpose2d(x2, wt2, b, stride=2, padding=1, output_padding=1)) < 1.5e-3


def test_avgpool_and_im2col_k3(dev):
    from spatiotemporalentropymodel_b200 import _lib
    from spatiotemporalentropymodel_b200.engine import ConvOp, avgpool_nhwc
    g = torch.Generator().manual_seed(10)
    B, C, h, w = 2, 128, 12, 20
    x = rnd(g, (B, C, h, w))
    out = avgpool_nhwc(to_nhwc16(x, dev), torch.empty((B, h // 2, w // 2, C), dtype=torch.float16, device=dev), 2)
    assert torch.allclose(to_nchw(out), F.adaptive_avg_pool2d(x, (h // 2, w // 2)), atol=2e-3)
    # first quality-map conv: conv(4, 192, 3, 1) on cat[x, Qmap] through im2col rows + 1x1 GEMM
    img, q = torch.rand((B, 3, h, w), generator=g), torch.rand((B, 1, h, w), generator=g)
    wt, b = rnd(g, (192, 4, 3, 3), 1 / 6.0), 0.1 * torch.randn(192, generator=g)
    rows = torch.empty((B, h, w, 40), dtype=torch.float16, device=dev)
    img_d, q_d = img.to(dev), q.to(dev)  # keep the device copies alive (temporaries would alias in the allocator)
    _lib.check(_lib.load().stemb200_im2col_k3s1_c4(img_d.data_ptr(), q_d.data_ptr(), rows.data_ptr(), B, h,
                                                   w, torch.cuda.current_stream().cuda_stream), "im2col_k3s1_c4")
    w40 = F.pad(wt.permute(0, 2, 3, 1).reshape(192, 36), (0, 4)).reshape(192, 40, 1, 1)
    op = ConvOp(w40.to(dev), b.to(dev), c_in=[40], c_out=192, k=1, slope=0.1)
    out = op([rows], B, h, w, torch.empty((B, h, w, 192), dtype=torch.float16, device=dev))
    ref = F.leaky_relu(F.conv2d(torch.cat([img, q], 1).half().float(), wt, b, padding=1), 0.1)
    assert rel_rms(to_nchw(out), ref) < 1.5e-3


# ------------------------------------------------------------------------------------------------------
# the same epilogues at sizes with >= 2 x 148 tiles: 2-CTA clusters (cta_group::2 MMAs, half a weight tile per CTA),
# every CTA pair walks several tiles (accumulator double buffering and the cross-CTA barriers wrap around)
# ------------------------------------------------------------------------------------------------------
def test_sft_epilogue_cluster_size(dev):
    from spatiotemporalentropymodel_b200.engine import sft_op
    g = torch.Generator().manual_seed(31)
    B, C, h, w, nh = 1, 128, 160, 256, 128
    x, a = rnd(g, (B, C, h, w)), rnd(g, (B, nh, h, w))
    wg, wb = rnd(g, (C, nh, 3, 3), 1 / math.sqrt(nh * 9)), rnd(g, (C, nh, 3, 3), 1 / math.sqrt(nh * 9))
    bg, bb = 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ref = F.leaky_relu(x * (1 + F.conv2d(a, wg, bg, padding=1)) + F.conv2d(a, wb, bb, padding=1), 0.2)
    op = sft_op(wg.to(dev), bg.to(dev), wb.to(dev), bb.to(dev), c_in=nh, slope=0.2)
    out = op([to_nhwc16(a, dev)], B, h, w, torch.full((B, h, w, C), float("nan"), dtype=torch.float16, device=dev),
             aux=to_nhwc16(x, dev))
    got = to_nchw(out)
    assert torch.isfinite(got).all() and rel_rms(got, ref) < 1.5e-3


@pytest.mark.parametrize("c_out", [192, 160])
def test_residual_and_odd_widths_cluster_size(dev, c_out):
    """residual epilogue (C_out = 192) and the 160-wide tile (80 weight rows per CTA) at cluster size"""
    from spatiotemporalentropymodel_b200.engine import ConvOp
    from spatiotemporalentropymodel_b200._lib import EPI_ADD
    g = torch.Generator().manual_seed(32 + c_out)
    B, C, h, w = 1, 192, 160, 256
    x, r = rnd(g, (B, C, h, w)), rnd(g, (B, c_out, h, w))
    wt, b = rnd(g, (c_out, C, 3, 3), 1 / math.sqrt(C * 9)), 0.1 * torch.randn(c_out, generator=g)
    out_buf = torch.full((B, h, w, c_out), float("nan"), dtype=torch.float16, device=dev)
    if c_out == 192:
        op = ConvOp(wt.to(dev), b.to(dev), c_in=[C], c_out=c_out, k=3, epilogue=EPI_ADD)
        out = op([to_nhwc16(x, dev)], B, h, w, out_buf, aux=to_nhwc16(r, dev))
        ref = F.conv2d(x, wt, b, padding=1) + r
    else:
        op = ConvOp(wt.to(dev), b.to(dev), c_in=[C], c_out=c_out, k=3, slope=0.1)
        out = op([to_nhwc16(x, dev)], B, h, w, out_buf)
        ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.1)
    got = to_nchw(out)
    assert torch.isfinite(got).all() and rel_rms(got, ref) < 1.5e-3
