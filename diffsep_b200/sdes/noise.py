"""Where the sampler's Gaussian noise comes from.

Default: drawn inside the fused update kernels (Philox4x32-10 keyed by ``(seed, per-draw offset)``,
``seed = torch.initial_seed()`` (+ the rank under torch.distributed) and the offset drawn from torch's default
generator, so ``torch.manual_seed`` controls and restarts it) — no noise tensor ever
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
        self._counter = 0        # draws made so far (diagnostics only)

    @staticmethod
    def _key():
        """Philox key: torch's seed, with the process rank mixed in under torch.distributed so that unseeded
        shards (all of which share torch's default seed) do not draw the same noise field."""
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                seed = (seed + 0x9E3779B97F4A7C15 * dist.get_rank()) & 0xFFFFFFFFFFFFFFFF
        except Exception:
            pass
        return seed

    def next(self, shape, device):
        """-> (tensor or None, seed, offset)."""
        if self._injected is not None:
            if not self._injected:
                raise RuntimeError("injected noise list exhausted")
            z = self._injected.pop(0)
            if tuple(z.shape) != tuple(shape):
                raise ValueError(f"injected noise has shape {tuple(z.shape)}, expected {tuple(shape)}")
            return z.to(device=device, dtype=torch.float32).contiguous(), 0, 0
        # the per-draw Philox offset comes from torch's default (CPU) generator, so torch.manual_seed(s) restarts
        # the noise stream — same seed, same samples, as with the reference's randn — without a counter to reset
        self._counter += 1
        off = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64).item())
        return None, self._key(), off


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
