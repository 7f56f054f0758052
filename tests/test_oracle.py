"""CPU tests: the oracle against the reference-generated golden fixtures and the reference's own KATs."""
import numpy as np
import pytest
import torch

from oracle import stem_oracle as O
from spatiotemporalentropymodel_b200 import synthetic as S


def t(a):
    return torch.from_numpy(np.asarray(a))


# ---------------------------------------------------------------- reference unit-test KATs
def test_gdn_closed_form_at_init():
    """compressai_tests/test_layers.py:118-143: y = x / sqrt(1 + .1 x^2), inverse y = x * sqrt(1 + .1 x^2)."""
    c = 8
    ped = torch.tensor([2.0 ** -36])
    beta_p = torch.sqrt(torch.max(torch.ones(c) + ped, ped))
    gamma_p = torch.sqrt(torch.max(0.1 * torch.eye(c) + ped, ped))
    x = torch.rand(1, c, 5, 7) * 4 - 2
    assert torch.allclose(O.gdn(x, beta_p, gamma_p, False), x / torch.sqrt(1 + 0.1 * x ** 2), atol=1e-6)
    assert torch.allclose(O.gdn(x, beta_p, gamma_p, True), x * torch.sqrt(1 + 0.1 * x ** 2), atol=1e-6)


def test_mask_a_pattern():
    """compressai_tests/test_layers.py:38-50"""
    sd = {"context_prediction.weight": torch.ones(4, 3, 5, 5)}
    w = O.masked_weight(sd)
    assert w[:, :, :2].eq(1).all() and w[:, :, 2, :2].eq(1).all()
    assert w[:, :, 2, 2:].eq(0).all() and w[:, :, 3:].eq(0).all()
    assert int(w[0, 0].sum()) == 12


def test_round_semantics():
    """compressai_tests/test_entropy_models.py:58-71,249-286"""
    x = torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 0.49, 1.7])
    assert O.quantize_symbols(x).tolist() == [-2, -2, 0, 0, 2, 2, 0, 2]
    m = torch.rand(8)
    assert torch.equal(O.quantize_dequantize(x, m), torch.round(x - m) + m)
    y_hat, _ = O.gaussian_conditional_forward(x, torch.ones(8), None)
    assert torch.equal(y_hat, torch.round(x))


def test_scale_table_ends():
    """compressai_tests/test_models.py:173-181"""
    tb = O.get_scale_table()
    assert len(tb) == 64 and abs(tb[0].item() - 0.11) < 1e-7 and abs(tb[-1].item() - 256) < 1e-3


# ---------------------------------------------------------------- golden fixtures from the reference classes
def test_gaussian_conditional_kat(golden):
    g = golden("gaussian_conditional_kat.npz")
    y, mu, sigma = t(g["y"]), t(g["mu"]), t(g["sigma"])
    y_hat, lik = O.gaussian_conditional_forward(y, sigma, mu)
    assert torch.equal(y_hat, t(g["y_hat"]))
    assert torch.equal(lik, t(g["lik"]))
    assert torch.equal(O.build_indexes(sigma, t(g["scale_table"])), t(g["idx"]))
    assert torch.equal(O.quantize_symbols(y, mu), t(g["sym"]))
    # the survey's hand-checked vector
    assert g["idx"][3000:3008].tolist() == [13, 0, 0, 18, 1, 30, 63, 1]
    assert g["sym"][3000:3008].tolist() == [1, -2, 0, 2, 0, 11, -1, 0]


def test_entropy_bottleneck_kat(golden):
    g = golden("entropy_bottleneck_kat.npz")
    sd = S.make_stem_state_dict("SpatioTemporalPriorModel", seed=0)
    z_hat, lik = O.entropy_bottleneck_forward(t(g["z"]), sd)
    assert torch.equal(z_hat, t(g["z_hat"]))
    assert torch.allclose(lik, t(g["lik"]), rtol=1e-6, atol=0)


def test_pmf_to_quantized_cdf_kat(golden):
    g = golden("pmf_to_quantized_cdf_kat.npz")
    n = len([k for k in g if k.startswith("pmf")])
    assert n >= 10
    for i in range(n):
        assert O.pmf_to_quantized_cdf(g[f"pmf{i}"], 16).tolist() == g[f"cdf{i}"].tolist(), i
    assert O.pmf_to_quantized_cdf([0.1, 0.2, 0.3, 0.4]).tolist() == [0, 6554, 19661, 39322, 65536]
    assert O.pmf_to_quantized_cdf([1e-9, 0.5, 0.5, 1e-9]).tolist() == [0, 1, 32767, 65535, 65536]


CAL_TAG = {"default": "", "lowrate": "lowrate_"}


@pytest.mark.parametrize("calibration", S.CALIBRATIONS)
@pytest.mark.parametrize("variant", S.STEM_VARIANTS)
def test_stem_forward_matches_reference(golden, variant, calibration):
    g = golden(f"stem_{CAL_TAG[calibration]}{variant}.npz")
    sd = S.make_stem_state_dict(variant, seed=0, calibration=calibration)
    out = O.stem_forward(variant, t(g["y_cur"]), t(g["y_cond"]), sd)
    assert torch.equal(out["y_hat"], t(g["y_hat"]))
    assert torch.allclose(out["likelihoods"]["y"], t(g["lik_y"]), rtol=1e-5, atol=0)
    assert torch.allclose(out["likelihoods"]["z"], t(g["lik_z"]), rtol=1e-5, atol=0)


@pytest.mark.parametrize("calibration", S.CALIBRATIONS)
def test_transforms_match_reference(golden, calibration):
    g = golden(f"stem_{CAL_TAG[calibration]}SpatioTemporalPriorModel.npz")
    sd_i = S.make_iframe_state_dict(seed=0, calibration=calibration)
    lo, hi = (0.0, 1.0) if calibration == "default" else S.LOWRATE_FRAME_RANGE
    frames = S.make_frames(2, 256, 256, seed=1234, lo=lo, hi=hi)
    y = O.g_a(frames[1:2], sd_i)
    assert torch.allclose(y, t(g["y_cur"]), rtol=1e-5, atol=2e-5)
    y0 = O.g_a(frames[0:1], sd_i)
    ties = (y0 - torch.floor(y0) - 0.5).abs() < 1e-4     # a latent within 1e-4 of a rounding tie may flip across builds
    assert torch.equal(torch.round(y0)[~ties], t(g["y_cond"])[~ties]) and float(ties.float().mean()) < 1e-3
    x_hat = O.g_s(t(g["y_hat"]), sd_i)
    assert torch.allclose(x_hat, t(g["x_hat"]), rtol=1e-5, atol=1e-5)
    if calibration == "lowrate":
        mse = float(((frames[1:2] - x_hat) ** 2).mean())
        assert abs(-10 * np.log10(mse) - float(g["psnr"])) < 1e-3


def test_lowrate_calibration_operating_point(golden):
    """The second synthetic checkpoint must sit where the parity gates can see errors:
    < 1 % of the y likelihoods on the 1e-9 floor (and < 5 % of the bits from floored elements), < 1 % of the
    reconstructed pixels clamped, PSNR in the range of a trained codec, |y - mu| of the order of sigma."""
    for variant in S.STEM_VARIANTS:
        g = golden(f"stem_lowrate_{variant}.npz")
        ly = t(g["lik_y"]).double()
        floored = ly <= 1.0001e-9
        bits = -torch.log2(ly)
        assert float(floored.float().mean()) < 0.01, variant
        assert float(bits[floored].sum() / bits.sum()) < 0.05, variant
        assert 20.0 < float(g["psnr"]) < 40.0, variant
        sd = S.make_stem_state_dict(variant, seed=0, calibration="lowrate")
        out = O.stem_forward(variant, t(g["y_cur"]), t(g["y_cond"]), sd, return_params=True)
        target = t(g["y_cur"]) - t(g["y_cond"]) if variant.endswith("_Res") else t(g["y_cur"])
        r = (target - out["means"]) / out["scales"].clamp_min(0.11)
        assert 0.5 < float(r.std()) < 1.5, variant
        assert float(out["scales"].std() / out["scales"].mean()) > 0.3, variant   # sigma really varies
    g = golden("stem_lowrate_SpatioTemporalPriorModel.npz")
    x_hat = t(g["x_hat"])
    assert float(((x_hat <= 0) | (x_hat >= 1)).float().mean()) < 0.01


def test_cdf_tables_match_reference(golden):
    g = golden("stem_SpatioTemporalPriorModel.npz")
    cdf, offset, length = O.gaussian_conditional_tables()
    assert torch.equal(cdf, t(g["gc_quantized_cdf"]))
    assert torch.equal(offset, t(g["gc_offset"]))
    assert torch.equal(length, t(g["gc_cdf_length"]))
    assert cdf.shape == (64, 3133) and cdf[0, :5].tolist() == [0, 1, 65534, 65535, 65536]


def test_pframe_forward_shapes_and_padding():
    """evalSTEM.py:96-109: 1080 rows -> 4 + 4 padding; here a small non-multiple-of-64 frame."""
    sd_i = S.make_iframe_state_dict(seed=0)
    sd_s = S.make_stem_state_dict("SpatioTemporalPriorModel", seed=0)
    x = S.make_frames(1, 72, 100, seed=5)
    xp, (l, r, tp, b) = O.pad_to_64(x)
    assert xp.shape[-2:] == (128, 128) and (l, r, tp, b) == (14, 14, 28, 28)
    out = O.pframe_forward(x, S.make_latent(1, 192, 8, 8), sd_i, sd_s, "SpatioTemporalPriorModel")
    assert out["x_hat"].shape == x.shape and out["bpp"].shape == (1,) and torch.isfinite(out["psnr"]).all()
