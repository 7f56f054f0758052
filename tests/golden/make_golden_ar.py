"""Golden fixtures for the autoregressive coding rows, produced by the REFERENCE classes
(CPU): the I-frame model's forward / compress / decompress (priors.py:477-684) and the AR compress / decompress of
the STEM variants with a spatial context model (spatiotemporalpriors.py:588-768, :871-1055).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_ar.py
Inputs of the STEM cases are the latents already stored in tests/golden/stem_<variant>.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, prepare_reference  # noqa: E402


def main():
    prepare_reference()
    import warnings
    warnings.filterwarnings("ignore")
    from compressai.models import spatiotemporalpriors as ref_stem
    from compressai.zoo import models as ref_models
    from spatiotemporalentropymodel_b200 import synthetic as S

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)
    sd_i = S.make_iframe_state_dict(seed=0)
    iframe = ref_models["mbt2018"](quality=4)
    iframe.load_state_dict(sd_i)
    iframe.update(force=True)
    iframe.eval()
    rec = {}
    with torch.no_grad():
        x = S.make_frames(1, 128, 192, seed=77)          # latent 8 x 12, z 2 x 3
        out = iframe(x)
        rec.update(x=x.numpy(), y=out["y"].numpy(), y_hat=out["y_hat"].numpy(), x_hat=out["x_hat"].numpy(),
                   lik_y=out["likelihoods"]["y"].numpy(), lik_z=out["likelihoods"]["z"].numpy(),
                   scales_hat=out["entropy_params"]["scales_hat"].numpy(),
                   means_hat=out["entropy_params"]["means_hat"].numpy())
        enc = iframe.compress(x)
        dec = iframe.decompress(enc["strings"], enc["shape"])
        rec.update(y_string=np.frombuffer(enc["strings"][0][0], dtype=np.uint8),
                   z_string=np.frombuffer(enc["strings"][1][0], dtype=np.uint8),
                   shape=np.asarray(list(enc["shape"])), dec_y_hat=dec["y_hat"].numpy(), dec_x_hat=dec["x_hat"].numpy())
        bits = float((-torch.log2(out["likelihoods"]["y"])).sum() + (-torch.log2(out["likelihoods"]["z"])).sum())
        print(f"iframe: est bits {bits:.0f}, coded bits {8 * (len(enc['strings'][0][0]) + len(enc['strings'][1][0]))}")
    np.savez_compressed(os.path.join(OUT, "iframe_codec.npz"), **rec)

    for variant in ("SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res", "SpatioTemporalPriorModelWithoutTPM"):
        g = np.load(os.path.join(OUT, f"stem_{variant}.npz"))
        y_cur, y_cond = torch.from_numpy(g["y_cur"]), torch.from_numpy(g["y_cond"])
        stem = getattr(ref_stem, variant)()
        stem.load_state_dict(S.make_stem_state_dict(variant, seed=0))
        stem.update(force=True)
        stem.eval()
        with torch.no_grad():
            enc = stem.compress(y_cur, y_cond)
            dec = stem.decompress(enc["strings"], enc["shape"], y_cond)
        y_hat = dec["y_hat"] if isinstance(dec, dict) else dec
        np.savez_compressed(os.path.join(OUT, f"stem_ar_{variant}.npz"),
                            y_string=np.frombuffer(enc["strings"][0][0], dtype=np.uint8),
                            z_string=np.frombuffer(enc["strings"][1][0], dtype=np.uint8),
                            shape=np.asarray(list(enc["shape"])), dec_y_hat=y_hat.numpy())
        print(f"{variant}: y {len(enc['strings'][0][0])} B, z {len(enc['strings'][1][0])} B, "
              f"latent {tuple(y_cur.shape)}")


if __name__ == "__main__":
    main()
