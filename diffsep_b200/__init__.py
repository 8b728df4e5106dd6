"""diffsep_b200 — B200-native reverse-diffusion inference path of DiffSep
(fakufaku/diffusion-separation): STFT-510, NCSN++ score network, MixSDE predictor-corrector
sampler, behind the reference's own Python surfaces.  All arithmetic runs in ``libdsep.so``
(hand-written sm_100a kernels, C-ABI in ``include/dsep.h``); there is no CPU fallback."""
__all__ = ["DiffSepModel", "ScoreModelNCSNpp", "sdes", "ops"]


def __getattr__(name):
    import importlib
    if name in ("sdes", "ops", "backbone", "score_model", "pl_model", "build", "_lib"):
        return importlib.import_module(f".{name}", __name__)
    if name == "DiffSepModel":
        return importlib.import_module(".pl_model", __name__).DiffSepModel
    if name == "ScoreModelNCSNpp":
        return importlib.import_module(".score_model", __name__).ScoreModelNCSNpp
    raise AttributeError(name)
