"""Generate tests/golden/*.npz by running the REFERENCE classes (mmSir/SpatioTemporalEntropyModel) themselves.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It copies the reference tree to a scratch directory (the reference is read-only and its setup.py rewrites
version.py), builds the two pybind11 extensions there, adds the missing `compressai/models/gain.py`
(SURVEY.md §8c shim 2), imports the reference model classes, loads the seeded synthetic checkpoints of
`spatiotemporalentropymodel_b200.synthetic` into them with their own `load_state_dict` (strict, so the key /
shape contract is verified), and records their CPU outputs.  The fixtures pin both the oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py).
"""
import os
import shutil
import subprocess
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(REPO, "tests", "golden")
SCRATCH = os.environ.get("STEM_REF_SCRATCH", "/tmp/stemref")
sys.path.insert(0, REPO)


def prepare_reference():
    if not os.path.exists(os.path.join(SCRATCH, "compressai", "models", "gain.py")):
        if os.path.exists(SCRATCH):
            shutil.rmtree(SCRATCH)
        shutil.copytree("/root/reference", SCRATCH)
        for f in os.listdir(os.path.join(SCRATCH, "compressai")):
            if f.endswith(".pyd"):
                os.remove(os.path.join(SCRATCH, "compressai", f))
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        open(os.path.join(SCRATCH, "compressai", "models", "gain.py"), "w").close()
    sys.path.insert(0, SCRATCH)


def main():
    prepare_reference()
    import compressai  # noqa: F401  (the reference)
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.models import spatiotemporalpriors as ref_stem
    from compressai.zoo import models as ref_models
    from compressai._CXX import pmf_to_quantized_cdf as ref_pmf_to_cdf

    from spatiotemporalentropymodel_b200 import synthetic as S

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)

    # ---------------------------------------------------------------- I-frame transforms + STEM variants
    # two synthetic checkpoints (synthetic.py): "default" -> stem_<variant>.npz, "lowrate" -> stem_lowrate_<variant>.npz
    for calibration in S.CALIBRATIONS:
        sd_i = S.make_iframe_state_dict(seed=0, calibration=calibration)
        iframe = ref_models["mbt2018"](quality=4)
        iframe.load_state_dict(sd_i)
        iframe.update(force=True)
        iframe.eval()
        tag = "" if calibration == "default" else f"{calibration}_"
        rng = (0.0, 1.0) if calibration == "default" else S.LOWRATE_FRAME_RANGE
        with torch.no_grad():
            for variant in S.STEM_VARIANTS:
                size = 256 if variant == "SpatioTemporalPriorModel" else 128
                frames = S.make_frames(2, size, size, seed=1234, lo=rng[0], hi=rng[1])
                y0, _ = iframe.getY(frames[0:1])
                y_cond = torch.round(y0)  # stand-in for the I-frame codec's decoded latent
                y_cur, _ = iframe.getY(frames[1:2])
                sd_s = S.make_stem_state_dict(variant, seed=0, calibration=calibration)
                stem = getattr(ref_stem, variant)()
                stem.load_state_dict(sd_s)
                stem.update(force=True)
                stem.eval()
                out = stem(y_cur, y_cond)
                rec = {
                    "y_cur": y_cur.numpy(), "y_cond": y_cond.numpy(), "y_hat": out["y_hat"].numpy(),
                    "lik_y": out["likelihoods"]["y"].numpy(), "lik_z": out["likelihoods"]["z"].numpy(),
                }
                x_hat = iframe.getX(out["y_hat"])
                if variant == "SpatioTemporalPriorModel":
                    rec["x_hat"] = x_hat.numpy()
                if variant == "SpatioTemporalPriorModel" and calibration == "default":
                    rec["gc_quantized_cdf"] = stem.gaussian_conditional._quantized_cdf.numpy()
                    rec["gc_offset"] = stem.gaussian_conditional._offset.numpy()
                    rec["gc_cdf_length"] = stem.gaussian_conditional._cdf_length.numpy()
                    rec["eb_quantized_cdf"] = stem.entropy_bottleneck._quantized_cdf.numpy()
                    rec["eb_offset"] = stem.entropy_bottleneck._offset.numpy()
                    rec["eb_cdf_length"] = stem.entropy_bottleneck._cdf_length.numpy()
                if calibration != "default":
                    # evalSTEM.py:127-136 statistics of this frame, as the reference computes them
                    mse = float(((frames[1:2] - x_hat) ** 2).mean())
                    rec["psnr"] = np.float64(-10 * np.log10(mse))
                np.savez_compressed(os.path.join(OUT, f"stem_{tag}{variant}.npz"), **rec)
                ly, lz = out["likelihoods"]["y"], out["likelihoods"]["z"]
                bits = float((-torch.log2(ly)).sum() + (-torch.log2(lz)).sum())
                mse = float(((frames[1:2] - x_hat) ** 2).mean())
                print(f"[{calibration}] {variant}: size {size}, bpp {bits / (size * size):.4f}, PSNR "
                      f"{-10 * np.log10(mse):.2f} dB, floored y likelihoods {float((ly <= 1.0001e-9).float().mean()):.4f}, "
                      f"clamped x_hat pixels {float(((x_hat <= 0) | (x_hat >= 1)).float().mean()):.4f}")

    # ---------------------------------------------------------------- stem_roi (a13), both calibrations
    from compressai.models.stem_roi import stem_roi as ref_stem_roi
    from spatiotemporalentropymodel_b200 import stem_roi as R
    for calibration in ("default", "lowrate"):
        sd_r = R.make_synthetic_state_dict(seed=0, calibration=calibration)
        roi = ref_stem_roi()
        roi.load_state_dict(sd_r)
        roi.update(force=True)
        roi.eval()
        frames = S.make_frames(2, 128, 128, seed=11)
        rec = {}
        with torch.no_grad():
            for name, qmap in (("ramp", R.make_qmap(1, 128, 128, "ramp")), ("uniform", R.make_qmap(1, 128, 128, "uniform", 0.25))):
                out = roi(frames[1:2], frames[0:1], qmap)
                rec[f"{name}_x_hat"] = out["x_hat"].numpy()
                rec[f"{name}_y_hat"] = out["y_hat"].numpy()
                rec[f"{name}_lik_y"] = out["likelihoods"]["y"].numpy()
                rec[f"{name}_lik_z"] = out["likelihoods"]["z"].numpy()
                ly = out["likelihoods"]["y"]
                bits = float((-torch.log2(ly)).sum() + (-torch.log2(out["likelihoods"]["z"])).sum())
                print(f"[{calibration}] stem_roi[{name}]: bpp {bits / (128 * 128):.4f}, floored y likelihoods "
                      f"{float((ly <= 1.0001e-9).float().mean()):.4f}")
        np.savez_compressed(os.path.join(OUT, "stem_roi.npz" if calibration == "default" else "stem_roi_lowrate.npz"), **rec)

    # ---------------------------------------------------------------- isolated GaussianConditional (a9)
    table = ref_stem.get_scale_table()
    gc = GaussianConditional(None)
    gc.update_scale_table(table, force=True)
    gc.eval()
    g = torch.Generator().manual_seed(99)
    n = 8192
    y = 4.0 * torch.randn(n, generator=g)
    y[:512] = torch.round(y[:512]) + 0.5                       # exact .5 ties (half-to-even)
    mu = (torch.rand(n, generator=g) * 4 - 2)
    mu[:256] = 0.0
    sigma = torch.exp(torch.rand(n, generator=g) * (np.log(400.0) - np.log(0.02)) + np.log(0.02))
    edges = torch.cat([table, torch.nextafter(table, torch.tensor(float("inf"))),
                       torch.nextafter(table, torch.tensor(0.0))])
    sigma[1024:1024 + edges.numel()] = edges                   # every table entry and its neighbours
    sigma[2000:2008] = torch.tensor([0.0, -1.0, 0.11, 0.10999999, 1e-8, 256.0, 300.0, 1e4])
    # survey KAT
    y[3000:3008] = torch.tensor([1.7, -2.5, 0.5, 1.5, 0.04, 10.2, -0.49, 3.0])
    mu[3000:3008] = torch.tensor([0.3, 0, 0, 0, 0, -0.7, 0.02, 2.5])
    sigma[3000:3008] = torch.tensor([0.5, 0.05, 0.11, 1.0, 0.12, 4.0, 300, 0.1244])
    with torch.no_grad():
        y_hat, lik = gc(y, sigma, means=mu)
        idx = gc.build_indexes(sigma)
        sym = gc.quantize(y, "symbols", mu)
    np.savez_compressed(os.path.join(OUT, "gaussian_conditional_kat.npz"), y=y.numpy(), mu=mu.numpy(),
                        sigma=sigma.numpy(), y_hat=y_hat.numpy(), lik=lik.numpy(), idx=idx.numpy(),
                        sym=sym.numpy(), scale_table=table.numpy())

    # ---------------------------------------------------------------- EntropyBottleneck forward (a4)
    sd_s = S.make_stem_state_dict("SpatioTemporalPriorModel", seed=0)
    eb = EntropyBottleneck(256)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd_s.items()
                        if k.startswith("entropy_bottleneck.") and "_offset" not in k and "_quantized_cdf" not in k
                        and "_cdf_length" not in k}, strict=False)
    eb.eval()
    z = 5.0 * torch.randn((2, 256, 5, 7), generator=g)
    with torch.no_grad():
        z_hat, z_lik = eb(z)
    np.savez_compressed(os.path.join(OUT, "entropy_bottleneck_kat.npz"), z=z.numpy(), z_hat=z_hat.numpy(),
                        lik=z_lik.numpy())

    # ---------------------------------------------------------------- rANS byte streams (cpp_exts/rans)
    from compressai.ans import RansDecoder, RansEncoder
    gr = torch.Generator().manual_seed(31)
    n_sym = 20000
    ridx = torch.randint(0, 64, (n_sym,), generator=gr, dtype=torch.int32)
    # symbols drawn per scale, with heavy tails so that the bypass (escape) path is exercised
    rsym = torch.round(table[ridx.long()] * torch.randn(n_sym, generator=gr) * 1.5).int()
    rsym[::97] += 4000
    rsym[5::131] -= 7000
    cdfs = gc._quantized_cdf.tolist()
    sizes = gc._cdf_length.reshape(-1).int().tolist()
    offs = gc._offset.reshape(-1).int().tolist()
    stream = RansEncoder().encode_with_indexes(rsym.tolist(), ridx.tolist(), cdfs, sizes, offs)
    assert RansDecoder().decode_with_indexes(stream, ridx.tolist(), cdfs, sizes, offs) == rsym.tolist()
    np.savez_compressed(os.path.join(OUT, "rans_kat.npz"), symbols=rsym.numpy(), indexes=ridx.numpy(),
                        stream=np.frombuffer(stream, dtype=np.uint8))
    print(f"rANS KAT: {n_sym} symbols -> {len(stream)} bytes")

    # ---------------------------------------------------------------- pmf_to_quantized_cdf (ops.cpp)
    pmfs = [[0.1, 0.2, 0.3, 0.4], [1e-9, 0.5, 0.5, 1e-9], [0.25] * 4, [1e-12] * 6 + [1.0], [0.3, 1e-7, 0.7 - 1e-7]]
    gp = torch.Generator().manual_seed(7)
    for _ in range(6):
        p = torch.rand(int(torch.randint(3, 40, (1,), generator=gp)), generator=gp) ** 6
        pmfs.append((p / p.sum()).tolist())
    rec = {}
    for i, p in enumerate(pmfs):
        rec[f"pmf{i}"] = np.asarray(p, dtype=np.float32)
        rec[f"cdf{i}"] = np.asarray(ref_pmf_to_cdf([float(np.float32(v)) for v in p], 16), dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "pmf_to_quantized_cdf_kat.npz"), **rec)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
