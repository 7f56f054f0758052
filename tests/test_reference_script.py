"""The UNMODIFIED per-frame functions of the reference's evaluation script (stem/evalSTEM.py:34-154: inferenceI_DVR,
inferenceP_DVR) running on this package through ``compat.install()``.

The script itself is not part of this repository (reference sources are never copied) and /root/reference does not
exist on the GPU boxes, so the test is gated: set STEM_REFERENCE_SCRIPT to the path of the reference's evalSTEM.py
on a machine that has both the file and a B200.  The committed log of such a run is
profiles/r02_unmodified_evalSTEM_functions.log.  `pytorch_msssim` (imported by the script, absent from this image,
MS-SSIM is out of scope - SURVEY.md §8c) is stubbed; nothing else is touched: the module is imported from its file as
is, so its `from compressai.zoo import models` / `from compressai.models.spatiotemporalpriors import *` resolve to the
B200 classes, and its functions call getY / forward / compress / decompress / getX with the reference's signatures.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

from spatiotemporalentropymodel_b200 import synthetic as S

SCRIPT = os.environ.get("STEM_REFERENCE_SCRIPT", "/root/reference/stem/evalSTEM.py")


def load_reference_script():
    from spatiotemporalentropymodel_b200 import compat
    compat.install(force=True)
    if "pytorch_msssim" not in sys.modules:
        stub = types.ModuleType("pytorch_msssim")
        stub.ms_ssim = lambda a, b, data_range=1.0: torch.zeros(())   # MS-SSIM is out of scope
        sys.modules["pytorch_msssim"] = stub
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    spec = importlib.util.spec_from_file_location("ref_evalSTEM", SCRIPT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                                       # the script sets CUDA_VISIBLE_DEVICES = "0" (:26)
    if visible is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = visible
    return mod


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="the reference's stem/evalSTEM.py is not on this machine")
def test_unmodified_script_imports_resolve_to_b200_classes():
    """CPU part (runs in the build container): the script imports cleanly and every model entry point it calls exists
    with the reference's signature."""
    import ast
    import inspect
    from spatiotemporalentropymodel_b200 import models as M
    ev = load_reference_script()
    assert ev.models["mbt2018"] is M.mbt2018
    assert ev.SpatioTemporalPriorModel_Res is M.SpatioTemporalPriorModel_Res
    calls = {"IFrameCompressor": set(), "stem": set(), "model": set()}
    for fn in (ev.inferenceI_DVR, ev.inferenceP_DVR):
        for node in ast.walk(ast.parse(inspect.getsource(fn))):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and \
                    isinstance(node.func.value, ast.Name) and node.func.value.id in calls:
                calls[node.func.value.id].add(node.func.attr)
    assert calls["IFrameCompressor"] == {"getY", "getX"} and calls["stem"] == {"compress", "decompress"}
    assert calls["model"] == {"compress", "decompress"}
    net, stem = M.models["mbt2018"](quality=4), M.SpatioTemporalPriorModel_Res()
    for name in calls["IFrameCompressor"] | calls["model"]:
        assert callable(getattr(net, name))
    for name in calls["stem"]:
        assert callable(getattr(stem, name))
    assert list(inspect.signature(stem.compress).parameters) == ["y_cur", "y_conditioned"]
    assert list(inspect.signature(stem.decompress).parameters) == ["strings", "shape", "y_conditioned"]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="the reference's stem/evalSTEM.py is not on this machine")
@pytest.mark.parametrize("variant", ["SpatioTemporalPriorModel_Res", "SpatioTemporalPriorModelWithoutSPM"])
def test_unmodified_inference_functions_run_on_the_package(variant):
    """inferenceI_DVR on the first frame, inferenceP_DVR on the next two (evalSTEM.py:186-209 without the dataset
    paths and the CPU bounce), against this package's own evaluation driver on the same frames."""
    from spatiotemporalentropymodel_b200 import evaluate as E, models as M
    ev = load_reference_script()
    dev = torch.device("cuda:0")
    net = ev.models["mbt2018"](quality=4)
    net.load_state_dict(S.make_iframe_state_dict(0))
    net.update(force=True)
    stem = getattr(ev, variant)()
    stem.load_state_dict(S.make_stem_state_dict(variant, 0))
    stem.update(force=True)
    net, stem = net.to(dev).eval(), stem.to(dev).eval()
    frames = S.make_frames(3, 120, 200, seed=9).to(dev)
    out = ev.inferenceI_DVR(net, frames[0])
    rows = [(out["bpp"], out["psnr"], out["estimate_bpp"])]
    y_cond = out["y_conditioned"]
    for t in (1, 2):
        out = ev.inferenceP_DVR(net, stem, frames[t], y_cond)
        y_cond = out["y_conditioned"]
        rows.append((out["bpp"], out["psnr"], out["estimate_bpp"]))
    want = E.code_gop(net, stem, frames, mode="real")
    for (bpp, psnr, est), (wb, wp) in zip(rows, want):
        assert abs(bpp - wb) < 1e-9 and abs(psnr - wp) < 1e-6, (rows, want)
        assert 0.5 * bpp < est < 2.0 * bpp        # estimate and coded size agree in scale (16 % floored likelihoods)
    print(variant, [tuple(round(v, 4) for v in r) for r in rows])
