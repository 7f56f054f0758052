"""Autoregressive coding rows: compress / decompress of the STEM variants with a spatial
context model and of the I-frame model (mbt2018).

CPU (-m "not gpu"): the oracle's restatement of the reference's raster scans reproduces the REFERENCE's own outputs
(tests/golden/stem_ar_*.npz, iframe_codec.npz, written by tests/golden/make_golden_ar.py from the reference classes):
decoded latents bit-exact, and the oracle's symbols pushed through the C-ABI rANS coder give the reference's y string
byte for byte.
GPU (-m gpu): the wavefront encoder / in-kernel-rANS decoder round trip is lossless and self-consistent (the decoder
reproduces the encoder's y_hat bit for bit), and agrees with the reference within what fp16 tensor-core priors
allow: the AR recurrence amplifies any 1-ulp difference in (sigma, mu) into different symbols further down the scan,
so streams cannot be byte-identical across implementations; bits are within 1 % and decoded latents within
rounding of the reference at >= 95 % of the elements (97.5 % measured)."""
import numpy as np
import pytest
import torch

from oracle import stem_oracle as O
from spatiotemporalentropymodel_b200 import synthetic as S
from spatiotemporalentropymodel_b200.entropy_models import GaussianConditional, rans_encode
from spatiotemporalentropymodel_b200.models import get_scale_table

AR_VARIANTS = ("SpatioTemporalPriorModel_Res", "SpatioTemporalPriorModelWithoutTPM", "SpatioTemporalPriorModel")


def _gc():
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table(), force=True)
    return gc.eval()


@pytest.mark.parametrize("variant", AR_VARIANTS[:2])  # the 16x16 full model takes ~1 min on CPU: covered on the GPU
def test_oracle_ar_scan_reproduces_reference(golden, variant):
    g, ga = golden(f"stem_{variant}.npz"), golden(f"stem_ar_{variant}.npz")
    sd = S.make_stem_state_dict(variant, seed=0)
    with torch.no_grad():
        out = O.stem_ar_code(variant, torch.from_numpy(g["y_cur"]), torch.from_numpy(g["y_cond"]), sd)
    assert np.array_equal(out["y_hat"].numpy(), ga["dec_y_hat"])
    gc = _gc()
    stream = rans_encode(out["symbols"], out["indexes"], gc._quantized_cdf, gc._cdf_length, gc._offset)
    assert stream == ga["y_string"].tobytes()
    # decoding direction of the scan: symbols -> the same latent
    has_tpm = variant != "SpatioTemporalPriorModelWithoutTPM"
    y_cond = torch.from_numpy(g["y_cond"])
    with torch.no_grad():
        parts = ([O.TPM(y_cond, sd)] if has_tpm else []) + [O.HD(out["z_hat"], sd)]
        t_hat, _, _ = O.ar_scan(None, torch.cat(parts, 1), sd, symbols=out["symbols"])
    y_hat = t_hat + y_cond if variant.endswith("_Res") else t_hat
    assert np.array_equal(y_hat.numpy(), ga["dec_y_hat"])


def test_oracle_iframe_reproduces_reference(golden):
    g = golden("iframe_codec.npz")
    sd = S.make_iframe_state_dict(seed=0)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        f = O.iframe_forward(x, sd)
        c = O.iframe_ar_code(x, sd)
    assert np.array_equal(f["y_hat"].numpy(), g["y_hat"])
    np.testing.assert_allclose(f["lik_y"].numpy(), g["lik_y"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(f["lik_z"].numpy(), g["lik_z"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(f["x_hat"].numpy(), g["x_hat"], rtol=0, atol=1e-5)
    assert np.array_equal(c["y_hat"].numpy(), g["dec_y_hat"])
    gc = _gc()
    assert rans_encode(c["symbols"], c["indexes"], gc._quantized_cdf, gc._cdf_length,
                       gc._offset) == g["y_string"].tobytes()


# ------------------------------------------------------------------------------------------------------
def _stem(variant, dev):
    from spatiotemporalentropymodel_b200 import models as M
    m = getattr(M, variant)()
    m.load_state_dict(S.make_stem_state_dict(variant, seed=0))
    m.update(force=True)
    return m.to(dev).eval()


def _agreement(y_hat, ref):
    """fraction of latent elements whose decoded value is the reference's up to float noise in mu"""
    return float((np.abs(y_hat - ref) < 0.25).mean())


@pytest.mark.gpu
@pytest.mark.parametrize("variant", AR_VARIANTS)
def test_stem_ar_round_trip_gpu(golden, variant):
    dev = torch.device("cuda:0")
    g, ga = golden(f"stem_{variant}.npz"), golden(f"stem_ar_{variant}.npz")
    m = _stem(variant, dev)
    y_cur, y_cond = torch.from_numpy(g["y_cur"]).to(dev), torch.from_numpy(g["y_cond"]).to(dev)
    enc = m.compress(y_cur, y_cond)
    assert list(enc["shape"]) == list(ga["shape"])
    assert len(enc["strings"][0]) == 1 and isinstance(enc["strings"][0][0], bytes)
    dec = m.decompress(enc["strings"], enc["shape"], y_cond)
    y_hat = dec["y_hat"]
    assert y_hat.shape == y_cur.shape
    # lossless: decoded symbols are the encoder's -> |y - y_hat| <= 1/2 everywhere (+ float slack)
    assert float((y_hat - y_cur).abs().max()) <= 0.5 + 1e-3
    # the encoder's own reconstruction (what the next frame is conditioned on) is what the decoder produces
    eng = m.engine()
    from spatiotemporalentropymodel_b200.engine import nchw_to_nhwc_f16
    B, C, h, w = y_cur.shape
    cond16 = nchw_to_nhwc_f16(y_cond, torch.empty((B, h, w, C), dtype=torch.float16, device=dev))
    z_hat = m.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"]).to(dev)
    zhat16 = nchw_to_nhwc_f16(z_hat.contiguous(), torch.empty((B, h // 4, w // 4, eng.zc), dtype=torch.float16, device=dev))
    target = (y_cur - y_cond if variant.endswith("_Res") else y_cur).permute(0, 2, 3, 1).contiguous()
    t_hat, sym, idx, params = eng.ar_head().encode(target, eng.static_priors(zhat16, cond16, B, h, w), eng.scale_table)
    enc_y_hat = t_hat.permute(0, 3, 1, 2) + (y_cond if variant.endswith("_Res") else 0)
    assert torch.equal(enc_y_hat, y_hat)
    assert torch.equal(params.permute(0, 3, 1, 2)[:, C:], dec["entropy_params"]["means_hat"])
    # against the reference's own coder
    n_ref, n = ga["y_string"].size, len(enc["strings"][0][0])
    assert abs(n - n_ref) / n_ref < 0.01, (n, n_ref)
    assert _agreement(y_hat.cpu().numpy(), ga["dec_y_hat"]) > 0.95
    # batch of two: every image is an independent stream, identical to coding it alone
    enc2 = m.compress(torch.cat([y_cur, y_cur.flip(3)]), torch.cat([y_cond, y_cond.flip(3)]))
    assert enc2["strings"][0][0] == enc["strings"][0][0]
    dec2 = m.decompress(enc2["strings"], enc2["shape"], torch.cat([y_cond, y_cond.flip(3)]))
    assert torch.equal(dec2["y_hat"][0:1], y_hat)
    assert float((dec2["y_hat"][1:2] - y_cur.flip(3)).abs().max()) <= 0.5 + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("variant", AR_VARIANTS[:2])
def test_ar_encode_with_oracle_priors_isolates_the_head(golden, variant):
    """Where does the residual disagreement with the reference's scan come from (missing #2)?  The AR kernel
    gets the ORACLE's fp32 prior term e0 = W0[:, priors] . cat(TPM(y_cond), HD(z_hat)) + b0 instead of the fp16
    tensor-core one, so the only differences left are the fp32 summation order of context conv + head inside the
    kernel vs ATen's.  Compared with the oracle's raster scan (bit-exact to the reference, test above): symbols and
    indexes agree on (almost) every element; what the product path loses beyond that is the fp16 operand rounding of the
    priors, amplified by the recurrence."""
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    g = golden(f"stem_{variant}.npz")
    sd = S.make_stem_state_dict(variant, seed=0)
    y_cur, y_cond = torch.from_numpy(g["y_cur"]), torch.from_numpy(g["y_cond"])
    has_tpm = variant != "SpatioTemporalPriorModelWithoutTPM"
    res = variant.endswith("_Res")
    with torch.no_grad():
        ref = O.stem_ar_code(variant, y_cur, y_cond, sd)
        parts = ([O.TPM(y_cond, sd)] if has_tpm else []) + [O.HD(ref["z_hat"], sd)]
        pri = torch.cat(parts, 1)
        n_static = pri.shape[1]
        w0 = sd["EPM.0.weight"][:, :n_static]
        e0 = F.conv2d(pri, w0, sd["EPM.0.bias"]).permute(0, 2, 3, 1).contiguous()
    m = _stem(variant, dev)
    eng = m.engine()
    target = (y_cur - y_cond if res else y_cur).permute(0, 2, 3, 1).contiguous().to(dev)
    t_hat, sym, idx, _ = eng.ar_head().encode(target, [], eng.scale_table, e0=e0.to(dev))
    B, C, h, w = y_cur.shape
    sym_ref = ref["symbols"].reshape(B, h, w, C)
    idx_ref = ref["indexes"].reshape(B, h, w, C)
    sym_ok = float((sym.cpu() == sym_ref).float().mean())
    idx_ok = float((idx.cpu() == idx_ref).float().mean())
    y_hat = t_hat.permute(0, 3, 1, 2).cpu() + (y_cond if res else 0)
    print(f"{variant}: oracle priors -> symbols equal {sym_ok:.5f}, indexes equal {idx_ok:.5f}, "
          f"max |y_hat - ref| {float((y_hat - ref['y_hat']).abs().max()):.3g}")
    assert sym_ok > 0.999 and idx_ok > 0.999, (sym_ok, idx_ok)
    # product path on the same input, for the record: fp16 priors
    enc = m.compress(y_cur.to(dev), y_cond.to(dev))
    dec = m.decompress(enc["strings"], enc["shape"], y_cond.to(dev))
    agree = _agreement(dec["y_hat"].cpu().numpy(), ref["y_hat"].numpy())
    print(f"{variant}: fp16 tensor-core priors -> decoded latents equal to the reference's at {agree:.4f}")


@pytest.mark.gpu
def test_ar_decode_rejects_corrupt_stream(golden):
    dev = torch.device("cuda:0")
    variant = "SpatioTemporalPriorModel_Res"
    g = golden(f"stem_{variant}.npz")
    m = _stem(variant, dev)
    y_cur, y_cond = torch.from_numpy(g["y_cur"]).to(dev), torch.from_numpy(g["y_cond"]).to(dev)
    enc = m.compress(y_cur, y_cond)
    with pytest.raises(ValueError):
        m.decompress([[enc["strings"][0][0][:-3]], enc["strings"][1]], enc["shape"], y_cond)   # not whole words
    bad = bytes(len(enc["strings"][0][0]))                                                      # all-zero stream
    out = m.decompress([[bad], enc["strings"][1]], enc["shape"], y_cond)                       # decodes to garbage,
    assert out["y_hat"].shape == y_cur.shape and bool(torch.isfinite(out["y_hat"]).all())      # never hangs / NaNs


@pytest.mark.gpu
def test_iframe_model_gpu(golden):
    """forward (priors.py:477-508), compress / decompress (:510-644) of mbt2018 q4 on the CUDA kernels."""
    from spatiotemporalentropymodel_b200 import models as M
    dev = torch.device("cuda:0")
    g = golden("iframe_codec.npz")
    net = M.models["mbt2018"](quality=4)
    net.load_state_dict(S.make_iframe_state_dict(seed=0))
    net.update(force=True)
    net = net.to(dev).eval()
    x = torch.from_numpy(g["x"]).to(dev)
    out = net(x)
    assert set(out) == {"y", "y_hat", "x_hat", "likelihoods", "entropy_params"}
    y_ref = g["y"]
    assert float(np.abs(out["y"].cpu().numpy() - y_ref).max()) < 2e-2 * max(1.0, float(np.abs(y_ref).max()))
    assert float((out["y_hat"].cpu().numpy() != g["y_hat"]).mean()) < 5e-3         # round(y) flips at .5 ties only
    bits = lambda ly, lz: float(-np.log2(ly.astype(np.float64)).sum() - np.log2(lz.astype(np.float64)).sum())
    b, b_ref = bits(out["likelihoods"]["y"].cpu().numpy(), out["likelihoods"]["z"].cpu().numpy()), bits(g["lik_y"], g["lik_z"])
    assert abs(b - b_ref) / b_ref < 5e-3, (b, b_ref)
    psnr = lambda a, b_: -10 * np.log10(((a - b_) ** 2).mean())
    xr = g["x"]
    assert float(((out["x_hat"].cpu().numpy() - g["x_hat"]) ** 2).mean()) < 5e-5    # a few round(y) flips
    assert abs(psnr(out["x_hat"].cpu().numpy(), xr) - psnr(g["x_hat"], xr)) < 0.01
    enc = net.compress(x)
    assert list(enc["shape"]) == list(g["shape"])
    dec = net.decompress(enc["strings"], enc["shape"])
    assert set(dec) == {"x_hat", "y_hat"}
    assert float((dec["y_hat"] - out["y"]).abs().max()) <= 0.5 + 2e-2
    n, n_ref = len(enc["strings"][0][0]), g["y_string"].size
    assert abs(n - n_ref) / n_ref < 0.01, (n, n_ref)
    assert _agreement(dec["y_hat"].cpu().numpy(), g["dec_y_hat"]) > 0.95
    assert abs(psnr(dec["x_hat"].cpu().numpy(), xr) - psnr(g["dec_x_hat"], xr)) < 0.02
    assert float(dec["x_hat"].min()) >= 0.0 and float(dec["x_hat"].max()) <= 1.0
