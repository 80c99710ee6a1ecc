"""CPU restatement of the reference's log-mel transform (Loader.transform_wav, cxai/utils/dataloading.py:138-176).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Parity status: PINNED -- the reference builds the transform from
``torchaudio.transforms.Spectrogram(n_fft, hop_length, power=None)`` and ``torchaudio.transforms.MelScale(n_mels,
n_stft, sample_rate)``; ``oracle/gen_golden_logmel.py`` runs exactly those torchaudio objects in the build container
and stores their outputs under ``tests/golden/logmel_*.npz``; ``tests/test_oracle_logmel.py`` checks this numpy fp64
restatement against them."""
from __future__ import annotations

import math

import numpy as np


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk'), evaluated in fp32 in torchaudio's order of
    operations: MelScale stores this fp32 matrix, so its rounding (~1e-5 relative on the filter slopes) is part of the
    reference's transform."""
    import torch
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes)).double().numpy()


def transform_wav(wav, sample_rate, n_fft, hop_length, n_mels, width, clamp=True, return_mel=False):
    """wav [B, n_samples] -> log-mel [B, 1, n_mels, width] (fp64)."""
    wav = np.asarray(wav, dtype=np.float64)
    B, N = wav.shape
    pad = n_fft // 2
    x = np.pad(wav, ((0, 0), (pad, pad)), mode="reflect")                       # torch.stft(center=True, pad_mode='reflect')
    n_frames = 1 + N // hop_length
    import torch
    window = torch.hann_window(n_fft).double().numpy()      # the fp32 periodic window torchaudio's Spectrogram stores
    idx = np.arange(n_frames)[:, None] * hop_length + np.arange(n_fft)[None, :]
    frames = x[:, idx] * window                                                 # [B, T, n_fft]
    spec = np.fft.rfft(frames, axis=-1)                                         # [B, T, F]
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels, sample_rate)
    mel = np.abs(spec) @ fb                                                     # [B, T, n_mels]
    logmel = np.log10(mel + 1e-7)
    if clamp:
        logmel = np.maximum(logmel, -4.0)
    out = np.transpose(logmel[:, 1:width + 1, :], (0, 2, 1))[:, None]
    if return_mel:
        return out, np.transpose(mel[:, 1:width + 1, :], (0, 2, 1))[:, None]
    return out


def synth_wav(B, n_samples, seed, sample_rate=16000):
    """Peak-normalised test signals: a few partials with vibrato plus noise bursts (deterministic)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / sample_rate
    out = np.zeros((B, n_samples))
    for b in range(B):
        for _ in range(4):
            f0 = rng.uniform(80, 3000); a = rng.uniform(0.2, 1.0)
            out[b] += a * np.sin(2 * np.pi * f0 * t + 3.0 * np.sin(2 * np.pi * rng.uniform(0.5, 6) * t))
        burst = rng.integers(0, n_samples - 2000)
        out[b, burst:burst + 2000] += rng.normal(0, 0.5, 2000)
        out[b] += rng.normal(0, 1e-3, n_samples)
        out[b] /= np.abs(out[b]).max()
    return out.astype(np.float32)
