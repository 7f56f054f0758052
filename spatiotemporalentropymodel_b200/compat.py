"""Make `import compressai...` resolve to this package for the names the STEM scripts use.

`stem/evalSTEM.py:23-24` does `from compressai.zoo import models` and
`from compressai.models.spatiotemporalpriors import *`; `:269-270,318` call
`compressai.available_entropy_coders()` / `set_entropy_coder()`.  `install()` registers alias modules in
`sys.modules` so those imports pick up the B200 classes without editing the script.  Call it before the script's
imports (and set the CUDA device first: the script overwrites CUDA_VISIBLE_DEVICES at import, evalSTEM.py:26).
"""
from __future__ import annotations

import sys
import types

_ENTROPY_CODER = "ans"


def available_entropy_coders():
    return ["ans"]


def get_entropy_coder():
    return _ENTROPY_CODER


def set_entropy_coder(entropy_coder: str) -> None:
    """compressai/__init__.py:33-48"""
    global _ENTROPY_CODER
    if entropy_coder not in available_entropy_coders():
        raise ValueError(f'Invalid entropy coder "{entropy_coder}", choose from'
                         f'({", ".join(available_entropy_coders())}).')
    _ENTROPY_CODER = entropy_coder


def install(force: bool = False) -> None:
    """Alias this package as `compressai` (no-op if a real compressai is already imported, unless force)."""
    if "compressai" in sys.modules and not force:
        return
    from . import entropy_models as em
    from . import models as m

    root = types.ModuleType("compressai")
    root.__version__ = "1.1.1+stemb200"
    root.available_entropy_coders = available_entropy_coders
    root.get_entropy_coder = get_entropy_coder
    root.set_entropy_coder = set_entropy_coder

    zoo = types.ModuleType("compressai.zoo")
    zoo.models = m.models
    zoo.mbt2018 = m.mbt2018

    models_pkg = types.ModuleType("compressai.models")
    stp = types.ModuleType("compressai.models.spatiotemporalpriors")
    names = ["SpatioTemporalPriorModel", "SpatioTemporalPriorModel_Res", "SpatioTemporalPriorModelWithoutSPM",
             "SpatioTemporalPriorModelWithoutTPM", "SpatioTemporalPriorModelWithoutSPMTPM", "get_scale_table"]
    for n in names:
        setattr(stp, n, getattr(m, n))
        setattr(models_pkg, n, getattr(m, n))
    stp.__all__ = names
    from . import stem_roi as roi
    roi_mod = types.ModuleType("compressai.models.stem_roi")
    roi_mod.stem_roi = roi.stem_roi
    models_pkg.stem_roi = roi_mod
    priors = types.ModuleType("compressai.models.priors")
    priors.CompressionModel = m.CompressionModel
    priors.JointAutoregressiveHierarchicalPriors = m.JointAutoregressiveHierarchicalPriors
    models_pkg.CompressionModel = m.CompressionModel
    models_pkg.JointAutoregressiveHierarchicalPriors = m.JointAutoregressiveHierarchicalPriors

    ent = types.ModuleType("compressai.entropy_models")
    ent.EntropyBottleneck = em.EntropyBottleneck
    ent.GaussianConditional = em.GaussianConditional
    ent.EntropyModel = em.EntropyModel

    layers = types.ModuleType("compressai.layers")
    layers.GDN = m.GDN
    layers.MaskedConv2d = m.MaskedConv2d

    root.zoo, root.models, root.entropy_models, root.layers = zoo, models_pkg, ent, layers
    models_pkg.spatiotemporalpriors, models_pkg.priors = stp, priors
    for name, mod in [("compressai", root), ("compressai.zoo", zoo), ("compressai.models", models_pkg),
                      ("compressai.models.spatiotemporalpriors", stp), ("compressai.models.priors", priors),
                      ("compressai.models.stem_roi", roi_mod),
                      ("compressai.entropy_models", ent), ("compressai.layers", layers)]:
        sys.modules[name] = mod
