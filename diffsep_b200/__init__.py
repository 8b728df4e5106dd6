"""diffsep_b200 — B200-native reverse-diffusion inference path of DiffSep
(fakufaku/diffusion-separation): STFT-510, NCSN++ score network, MixSDE predictor-corrector
sampler, behind the reference's own Python surfaces.  All arithmetic runs in ``libdsep.so``
(hand-written sm_100a kernels, C-ABI in ``include/dsep.h``); there is no CPU fallback."""
import os as _os

__all__ = ["DiffSepModel", "ScoreModelNCSNpp", "sdes", "ops", "DEFAULT_PASSES"]

# Tensor-core products per fp32-grade MAC of the convolutions: 2 = fp16 hi*hi + one e4m3 product
# carrying both correction terms (the product default), 3 = three fp16 products, 1 = hi*hi only (TF32-grade, not
# a parity mode).
DEFAULT_PASSES = int(_os.environ.get("DSEP_PASSES", "2"))


def __getattr__(name):
    import importlib
    if name in ("sdes", "ops", "backbone", "score_model", "pl_model", "build", "_lib"):
        return importlib.import_module(f".{name}", __name__)
    if name == "DiffSepModel":
        return importlib.import_module(".pl_model", __name__).DiffSepModel
    if name == "ScoreModelNCSNpp":
        return importlib.import_module(".score_model", __name__).ScoreModelNCSNpp
    raise AttributeError(name)
