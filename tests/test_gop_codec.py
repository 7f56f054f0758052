"""Real-bitstream GOP coding through the reference's evaluation loop (stem/evalSTEM.py:34-154, :186-209), restated with
this package's classes: I-frame compress -> decompress (mbt2018), then P-frames getY -> stem.compress ->
stem.decompress -> getX with y_conditioned = the decoded latent.  A second, independently constructed pair of models
(the "decoder", fed nothing but the strings, the shapes and its own previous output) must reproduce the encoder
side's reconstructions exactly - that is what makes the streams decodable."""
import math

import pytest
import torch
import torch.nn.functional as F

from spatiotemporalentropymodel_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _models(variant, dev):
    from spatiotemporalentropymodel_b200 import models as M
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(S.make_iframe_state_dict(seed=0))
    net.update(force=True)
    stem = getattr(M, variant)()
    stem.load_state_dict(S.make_stem_state_dict(variant, seed=0))
    stem.update(force=True)
    return net.to(dev).eval(), stem.to(dev).eval()


def _pad64(x):
    """evalSTEM.py:96-109"""
    h, w = x.size(2), x.size(3)
    nh, nw = (h + 63) // 64 * 64, (w + 63) // 64 * 64
    left, top = (nw - w) // 2, (nh - h) // 2
    pad = (left, nw - w - left, top, nh - h - top)
    return F.pad(x, pad, mode="constant", value=0), pad


def _psnr(a, b):
    return -10 * math.log10(float(F.mse_loss(a, b)))


@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel_Res", "SpatioTemporalPriorModel",
                                     "SpatioTemporalPriorModelWithoutSPM"])
def test_gop_real_bitstream_round_trip(variant):
    dev = torch.device("cuda:0")
    net, stem = _models(variant, dev)
    frames = S.make_frames(3, 120, 184, seed=21).to(dev)          # not a multiple of 64: exercises the padding
    packets, enc_side = [], []
    # ---- encoder (evalSTEM.inferenceI_DVR / inferenceP_DVR)
    x0, pad = _pad64(frames[0:1])
    enc = net.compress(x0)
    dec = net.decompress(enc["strings"], enc["shape"])
    packets.append(("I", enc["strings"], enc["shape"]))
    y_cond = dec["y_hat"]
    enc_side.append((y_cond, dec["x_hat"]))
    est_bits = real_bits = 0.0
    for t in (1, 2):
        xp, _ = _pad64(frames[t:t + 1])
        y_cur, _ = net.getY(xp)
        fwd = stem(y_cur, y_cond)
        enc = stem.compress(y_cur, y_cond)
        dec = stem.decompress(enc["strings"], enc["shape"], y_cond)
        assert set(dec) >= {"y_hat", "entropy_params"}
        y_cond = dec["y_hat"]
        x_hat = net.getX(y_cond)
        packets.append(("P", enc["strings"], enc["shape"]))
        enc_side.append((y_cond, x_hat))
        est_bits += sum(float(-torch.log2(v.double()).sum()) for v in fwd["likelihoods"].values())
        real_bits += 8.0 * sum(len(s[0]) for s in enc["strings"])
        assert float((y_cond - y_cur).abs().max()) <= 0.5 + 1e-3     # lossless symbols
    # the likelihood estimate floors at 1e-9 (29.9 bits) where the coder escapes for fewer: real <= estimate (+2 %)
    assert 0.5 * est_bits < real_bits < 1.02 * est_bits, (real_bits, est_bits)
    # ---- decoder: fresh models, strings only
    net2, stem2 = _models(variant, dev)
    y_prev = None
    for (kind, strings, shape), (y_ref, x_ref) in zip(packets, enc_side):
        if kind == "I":
            out = net2.decompress(strings, shape)
            y_prev, x_hat = out["y_hat"], out["x_hat"]
        else:
            y_prev = stem2.decompress(strings, shape, y_prev)["y_hat"]
            x_hat = net2.getX(y_prev)
        assert torch.equal(y_prev, y_ref)
        assert torch.equal(x_hat, x_ref)
    # reconstruction quality is that of the quantised latents (the synthetic checkpoints are not trained: only
    # consistency with the forward-mode reconstruction is checked)
    l, r, tp, b = pad
    for t, (_, x_hat) in enumerate(enc_side):
        crop = x_hat[:, :, tp:x_hat.size(2) - b, l:x_hat.size(3) - r]
        assert crop.shape == frames[t:t + 1].shape
        assert math.isfinite(_psnr(crop, frames[t:t + 1]))


def test_ar_batch_of_gops_medium_latent():
    """three independent images in one AR launch on a 32 x 48 latent (wavefronts of up to 16 positions x 3)"""
    dev = torch.device("cuda:0")
    _, stem = _models("SpatioTemporalPriorModel", dev)
    y_cond = torch.round(S.make_latent(3, 192, 32, 48, seed=5)).to(dev)
    y_cur = (y_cond + 0.7 * S.make_latent(3, 192, 32, 48, seed=6).to(dev)).contiguous()
    enc = stem.compress(y_cur, y_cond)
    assert len(enc["strings"][0]) == 3 and len(enc["strings"][1]) == 3
    dec = stem.decompress(enc["strings"], enc["shape"], y_cond)
    assert float((dec["y_hat"] - y_cur).abs().max()) <= 0.5 + 1e-3
    for b in range(3):                                           # every image decodes alone to the same latent
        one = stem.decompress([[enc["strings"][0][b]], [enc["strings"][1][b]]], enc["shape"], y_cond[b:b + 1])
        assert torch.equal(one["y_hat"], dec["y_hat"][b:b + 1])
