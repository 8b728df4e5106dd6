"""Where the sampler's Gaussian noise comes from.

Default: drawn inside the fused update kernels (Philox4x32-10 keyed by ``(seed, draw counter)``,
``seed = torch.initial_seed()`` so ``torch.manual_seed`` controls it) — no noise tensor ever
touches HBM.  For parity work, pre-drawn tensors can be injected in draw order (prior, then per
step ``[corrector] * n_steps``, predictor), which is how the tests compare against the reference
(seeds alone do not reproduce its ``randn_like`` on strided tensors).
"""
from __future__ import annotations

import contextlib

import torch


class NoiseSource:
    def __init__(self):
        self._injected = None
        self._counter = 0

    def next(self, shape, device):
        """-> (tensor or None, seed, offset)."""
        if self._injected is not None:
            if not self._injected:
                raise RuntimeError("injected noise list exhausted")
            z = self._injected.pop(0)
            if tuple(z.shape) != tuple(shape):
                raise ValueError(f"injected noise has shape {tuple(z.shape)}, expected {tuple(shape)}")
            return z.to(device=device, dtype=torch.float32).contiguous(), 0, 0
        self._counter += 1
        return None, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._counter


SOURCE = NoiseSource()


@contextlib.contextmanager
def injected_noise(tensors):
    """Within the context the sampler consumes ``tensors`` (a list, in draw order)."""
    prev = SOURCE._injected
    SOURCE._injected = list(tensors)
    try:
        yield SOURCE
    finally:
        SOURCE._injected = prev
