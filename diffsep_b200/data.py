"""Host-side batching helpers for folder-of-wav evaluation (the datasets of the reference are not
available here): ``max_collator`` keeps the reference's semantics (``datasets/wsj0_mix.py:95-111``:
pad every signal to the longest of the batch, centred: ``off // 2`` zeros in front)."""
from __future__ import annotations

import numpy as np
import torch


def max_collator(signals):
    """signals: list of [C, T_i] tensors -> ([B, C, T_max], [(front_pad, T_i), ...])."""
    max_len = max(s.shape[-1] for s in signals)
    out, spans = [], []
    for s in signals:
        off = max_len - s.shape[-1]
        out.append(torch.nn.functional.pad(s, (off // 2, off - off // 2)))
        spans.append((off // 2, s.shape[-1]))
    return torch.stack(out), spans


def wav_length(path):
    """number of samples per channel of a wav file, from its header (the data is memory-mapped, not read)"""
    from scipy.io import wavfile
    _, data = wavfile.read(path, mmap=True)
    return int(data.shape[0])


def bucket_by_length(lengths, batch_size):
    """Batches of indices in which every item has the SAME length (no padding), at most ``batch_size`` each, in
    order of first appearance.  The reference evaluates one utterance at a time (evaluate.py:340-376): normalisation,
    STFT framing and GroupNorm statistics see only that utterance, so a zero-padded ragged batch would change a short
    utterance's result; equal-length batches reproduce the batch-of-one results exactly."""
    order, groups = [], {}
    for i, n in enumerate(lengths):
        if n not in groups:
            groups[n] = []
            order.append(n)
        groups[n].append(i)
    batches = []
    for n in order:
        idx = groups[n]
        batches.extend(idx[k:k + batch_size] for k in range(0, len(idx), batch_size))
    return batches


def uncollate(batch, spans):
    """inverse of max_collator on the time axis: list of [..., T_i]"""
    return [batch[i, ..., front:front + n] for i, (front, n) in enumerate(spans)]


def load_wav(path):
    """-> (float32 tensor [channels, T] in [-1, 1), sample rate); scipy (torchaudio.load needs torchcodec)."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    elif data.dtype.kind == "u":
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = np.asarray(data, dtype=np.float32)
    data = data[None] if data.ndim == 1 else data.T
    return torch.from_numpy(np.ascontiguousarray(data)), sr


def save_wav(path, wav, sr):
    from scipy.io import wavfile
    wavfile.write(path, sr, wav.detach().cpu().numpy().T.astype(np.float32))
