"""smmd -- B200-native MMD^2 loss / KID score (drop-in for the hot path of playHing/Scaled-MMD-GAN).

    from smmd import mmd, compute_scores
    loss = mmd.mmd2(mmd._mix_rq_kernel(fake_features, real_features))      # fused fwd+bwd, autograd-aware
    kid = compute_scores.polynomial_mmd_averages(codes_g, codes_r, n_subsets=100, ret_var=False)

Importing the package does not load the CUDA library; the first call does and raises loudly if
``lib/libsmmd.so`` is missing or the device is not a B200 (there is no CPU / eager fallback).
"""
from . import _lib  # noqa: F401

__all__ = ["mmd", "compute_scores", "distributed"]


def __getattr__(name):
    if name in __all__:
        import importlib

        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
