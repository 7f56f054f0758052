"""CPU oracle for `stem_roi.forward` (compressai/models/stem_roi.py:585-608).  TEST INFRASTRUCTURE ONLY — same
rules as oracle/stem_oracle.py (never imported by the product package).  Functional restatement on a state_dict with
torch CPU ops; pinned by tests/golden/stem_roi.npz, which tests/golden/make_golden.py produces by running the
reference class itself on the seeded synthetic checkpoint."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import stem_oracle as O

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def _c(x, sd, name, stride=1, pad=None):
    w = sd[f"{name}.weight"]
    return F.conv2d(x, w, sd[f"{name}.bias"], stride=stride, padding=w.shape[-1] // 2 if pad is None else pad)


def _d(x, sd, name):
    w = sd[f"{name}.weight"]
    return F.conv_transpose2d(x, w, sd[f"{name}.bias"], stride=2, padding=w.shape[-1] // 2, output_padding=1)


def sft(x, qmap, sd, name):
    """stem_utils.py:36-43"""
    qmap = F.adaptive_avg_pool2d(qmap, x.size()[2:])
    actv = F.relu(_c(qmap, sd, f"{name}.mlp_shared.0"))
    return x * (1 + _c(actv, sd, f"{name}.mlp_gamma")) + _c(actv, sd, f"{name}.mlp_beta")


def sft_resblk(x, qmap, sd, name):
    """stem_utils.py:55-63"""
    dx = _c(F.leaky_relu(sft(x, qmap, sd, f"{name}.norm_0"), 0.2), sd, f"{name}.conv_0")
    dx = _c(F.leaky_relu(sft(dx, qmap, sd, f"{name}.norm_1"), 0.2), sd, f"{name}.conv_1")
    return x + dx


def _gdn(x, sd, name, inverse):
    return O.gdn(x, sd[f"{name}.beta"], sd[f"{name}.gamma"], inverse)


def _qfeat3(x, sd, name):
    """qmap_feature_{ga1,ha1,gs0}: three 3x3 stride-1 convs with LeakyReLU(0.1) between"""
    x = F.leaky_relu(_c(x, sd, f"{name}.0"), 0.1)
    x = F.leaky_relu(_c(x, sd, f"{name}.2"), 0.1)
    return _c(x, sd, f"{name}.4")


def _qfeat2(x, sd, name, transposed=False):
    x = _d(x, sd, f"{name}.0") if transposed else _c(x, sd, f"{name}.0", stride=2)
    return _c(F.leaky_relu(x, 0.1), sd, f"{name}.2")


def p_encoder(x, qmap, sd):
    """stem_roi.py:520-538"""
    q = _qfeat3(torch.cat([x, qmap], 1), sd, "qmap_feature_ga1")
    x = sft(_gdn(_c(x, sd, "ga1.0", stride=2), sd, "ga1.1", False), q, sd, "ga1_SFT")
    q = _qfeat2(q, sd, "qmap_feature_ga2")
    x = sft(_gdn(_c(x, sd, "ga2.0", stride=2), sd, "ga2.1", False), q, sd, "ga2_SFT")
    q = _qfeat2(q, sd, "qmap_feature_ga3")
    x = sft(_gdn(_c(x, sd, "ga3.0", stride=2), sd, "ga3.1", False), q, sd, "ga3_SFT")
    q = _qfeat2(q, sd, "qmap_feature_ga4")
    x = _c(x, sd, "ga4", stride=2)
    x = sft_resblk(x, q, sd, "ga4_SFTResB1")
    return sft_resblk(x, q, sd, "ga4_SFTResB2")


def condition_encoder(x, sd):
    """stem_roi.py:493-501"""
    for i in (0, 2, 4):
        x = _gdn(_c(x, sd, f"ConditionEncoder.{i}", stride=2), sd, f"ConditionEncoder.{i + 1}", False)
    return _c(x, sd, "ConditionEncoder.6", stride=2)


def hyper_encoder(x, qmap, sd):
    """stem_roi.py:562-579"""
    q = F.adaptive_avg_pool2d(qmap, x.size()[2:])
    q = _qfeat3(torch.cat([q, x], 1), sd, "qmap_feature_ha1")
    x = F.leaky_relu(sft(_c(x, sd, "ha1"), q, sd, "ha1_SFT"), 0.01)
    q = _qfeat2(q, sd, "qmap_feature_ha2")
    x = F.leaky_relu(sft(_c(x, sd, "ha2", stride=2), q, sd, "ha2_SFT"), 0.01)
    q = _qfeat2(q, sd, "qmap_feature_ha3")
    x = _c(x, sd, "ha3", stride=2)
    x = sft_resblk(x, q, sd, "ha3_ResB1")
    return sft_resblk(x, q, sd, "ha3_ResB2")


def p_decoder(y_hat, z_hat, sd):
    """stem_roi.py:540-560"""
    w = F.leaky_relu(_d(z_hat, sd, "wmap_generator.0"), 0.01)
    w = F.leaky_relu(_d(w, sd, "wmap_generator.2"), 0.01)
    w = _c(w, sd, "wmap_generator.4")
    w = _qfeat3(torch.cat([w, y_hat], 1), sd, "qmap_feature_gs0")
    x = sft_resblk(y_hat, w, sd, "gs0_SFTResB1")
    x = sft_resblk(x, w, sd, "gs0_SFTResB2")
    for i in (1, 2, 3):
        w = _qfeat2(w, sd, f"qmap_feature_gs{i}", transposed=True)
        x = sft(_gdn(_d(x, sd, f"gs{i}.0"), sd, f"gs{i}.1", True), w, sd, f"gs{i}_SFT")
    return _d(x, sd, "gs4")


def stem_roi_forward(x_cur: Tensor, x_cond: Tensor, qmap: Tensor, sd: SD):
    """stem_roi.py:585-608 (eval mode)"""
    y_cur = p_encoder(x_cur, qmap, sd)
    y_cond = condition_encoder(x_cond, sd)
    z = hyper_encoder(torch.cat([y_cur, y_cond], 1), qmap, sd)
    z_hat, z_lik = O.entropy_bottleneck_forward(z, sd)
    hp = O._seq_conv(z_hat, sd, "hs", [(0, "deconv", 2, 2), (2, "deconv", 2, 2), (4, "conv", 1, 1)])
    tp = O.TPM(y_cond, sd)
    gp = O.EPM(torch.cat([tp, hp], 1), sd)
    scales, means = gp.chunk(2, 1)
    y_hat, y_lik = O.gaussian_conditional_forward(y_cur, scales, means)
    x_hat = p_decoder(y_hat, z_hat, sd)
    return {"x_hat": x_hat, "y_hat": y_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "y_cur": y_cur, "z": z,
            "scales": scales, "means": means}
