"""Parity report of the CUDA P-frame pipeline against the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Used by ``tests/`` (the 1080p parity tests) and by ``bench.py``'s ``parity`` leg; the product package never imports
it.  It only *compares*: the caller runs ``PFramePipeline.forward_gop`` on the GPU and ``stem_oracle.gop_forward``
on the CPU with the same checkpoint and frames and hands both results over.

Gates (BASELINE.json north_star): per-frame bpp within 0.5 %, PSNR within 0.01 dB.  Everything else is reported so
that a failing gate can be traced: latent mismatch rate (rounding flips of y_hat), relative RMS of sigma and of mu
(the latter in units of sigma), share of floored likelihoods in the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

BPP_RTOL = 5e-3     # bpp within 0.5 %
PSNR_ATOL = 0.01    # PSNR within 0.01 dB


def _rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


def gop_parity(out: Dict[str, torch.Tensor], ref: List[Dict[str, torch.Tensor]], height: int, width: int,
               params_nhwc: Optional[torch.Tensor] = None) -> Dict[str, object]:
    """out: result of PFramePipeline.forward_gop / run_gop (device tensors); ref: list of per-frame dicts of
    stem_oracle.gop_forward(..., return_params=True) for the FIRST len(ref) frames of the GOP (frame t of a GOP only
    depends on frames <= t, so a prefix of a longer GPU run can be checked); params_nhwc: the engine's (T, h, w, 2C)
    sigma | mu buffer.  Returns {"ok", "frames": [...], "max_bpp_rel_err", "max_psnr_abs_err", ...}."""
    stats = out["stats"].detach().cpu().double()
    T = len(ref)
    assert 1 <= T <= stats.shape[1], (T, stats.shape)
    npx = height * width
    bpp = (stats[0] + stats[1]) / npx
    psnr = -10.0 * torch.log10(stats[2] / (3 * npx))
    y_hat = out["y_hat"].detach().cpu() if "y_hat" in out else None
    frames = []
    for t in range(T):
        r = ref[t]
        rb, rp = float(r["bpp"]), float(r["psnr"])
        rec = {"bpp": float(bpp[t]), "bpp_ref": rb, "bpp_rel_err": abs(float(bpp[t]) - rb) / rb,
               "psnr": float(psnr[t]), "psnr_ref": rp, "psnr_abs_err": abs(float(psnr[t]) - rp),
               "bpp_z_rel_err": abs(float(stats[1, t]) / npx - float(r["bpp_z"])) / max(float(r["bpp_z"]), 1e-12),
               "ref_floored_lik_frac": float((r["lik_y"] <= 1.0001e-9).float().mean())}
        if y_hat is not None:
            d = (y_hat[t:t + 1] - r["y_hat"]).abs()
            rec["y_hat_mismatch_frac"] = float((d > 1e-3).float().mean())
            rec["y_hat_max_abs_diff"] = float(d.max())
        if params_nhwc is not None and "scales" in r:
            p = params_nhwc[t:t + 1].detach().cpu().permute(0, 3, 1, 2)
            C = p.shape[1] // 2
            sg, mu = p[:, :C], p[:, C:]
            rs = r["scales"].clamp_min(0.11)
            rec["sigma_rel_rms"] = _rel_rms(sg.clamp_min(0.11), rs)
            rec["mu_rel_rms"] = _rel_rms(mu, r["means"])
            rec["mu_err_over_sigma_rms"] = float(torch.sqrt((((mu - r["means"]) / rs).double() ** 2).mean()))
        frames.append(rec)
    res = {
        "frames": frames,
        "max_bpp_rel_err": max(f["bpp_rel_err"] for f in frames),
        "max_psnr_abs_err": max(f["psnr_abs_err"] for f in frames),
        "gates": {"bpp_rel": BPP_RTOL, "psnr_db": PSNR_ATOL},
    }
    for k in ("y_hat_mismatch_frac", "sigma_rel_rms", "mu_err_over_sigma_rms"):
        if k in frames[0]:
            res["max_" + k] = max(f[k] for f in frames)
    res["ok"] = bool(res["max_bpp_rel_err"] <= BPP_RTOL and res["max_psnr_abs_err"] <= PSNR_ATOL
                     and all(math.isfinite(f["psnr"]) for f in frames))
    return res
