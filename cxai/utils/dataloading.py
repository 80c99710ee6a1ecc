"""Audio -> log-mel frontend -- mirror of the part of the reference's cxai/utils/dataloading.py that feeds the hot path:
``Loader`` (:13-176) with ``transform_wav`` (:138-176) running on the device (``logmel_transform_wav`` of
libdrsa_b200.so), and ``shuffle_and_truncate_databatch`` (:179-205).  File discovery helpers (fold lists, :208-320) are
outside the path.  ``Loader.load`` needs ``torchaudio`` for decoding only."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

from drsa_audio_b200 import _lib as _L
from cxai.utils.constants import AUDIO_PARAMS

__all__ = ["Loader", "shuffle_and_truncate_databatch", "get_slice", "peak_normalizer", "melscale_fbanks_htk"]


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(..., norm=None, mel_scale='htk') -> [n_freqs, n_mels], evaluated in fp32 in
    torchaudio's order of operations (the filter edges are rounded there, which moves the filters by ~1e-5 relative)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))


def peak_normalizer(wav: torch.Tensor) -> torch.Tensor:
    """Scale every waveform to a peak of 1."""
    return wav / wav.abs().amax(dim=-1, keepdim=True).clamp(min=1e-12)


def get_slice(wav: torch.Tensor, slice_length: int = 6, start_point: int = 0, num_chunks: int = 1,
              sample_rate: int = 16000) -> torch.Tensor:
    """Snippets of ``slice_length`` seconds from a waveform [channels, samples] (reference utils/sound.py:8-44).

    ``num_chunks > 1``: evenly spaced, OVERLAPPING windows over the first 29 s (the shortest GTZAN clip is ~29.3 s) with
    ``hop = floor((29 - slice_length) / (num_chunks - 1) * 10) / 10`` seconds, ``start_point`` ignored ->
    [channels * num_chunks, 1, window].  Otherwise one window starting at ``start_point`` seconds -> [channels, window]."""
    wav = torch.as_tensor(wav)
    window_size = int(slice_length * sample_rate)
    if num_chunks > 1:
        hop = int(math.floor(((29 - slice_length) / (num_chunks - 1)) * 10) / 10 * sample_rate)
        out = wav[:, :29 * sample_rate].unfold(1, window_size, hop).reshape(-1, 1, window_size)
        assert out.shape[0] == num_chunks, "not equal num_chunks"
        return out
    start_sample = int(start_point * sample_rate)
    assert start_point <= wav.size(1) - window_size, f"Start_point has to be in range [{0},{wav.size(1) - window_size}]"
    return wav[:, start_sample:start_sample + window_size]


class Loader:
    """Same constructor and methods as the reference class; the spectrogram transform runs on ``device``."""

    def __init__(self, case: str | None = None, sample_rate: int = 16000, n_fft: int = 800, hop_length: int = 360,
                 n_mels: int = 128, slice_length: int = 3, width: int = 128, device="cuda") -> None:
        if case is not None and case in list(AUDIO_PARAMS.keys()):
            p = AUDIO_PARAMS[case]
            self.sample_rate, n_fft, hop_length = p["sample_rate"], p["n_fft"], p["hop_length"]
            self.n_mels, self.width, self.slice_length = p["n_mels"], p["mel_width"], p.get("slice_length", 0)
        else:
            self.sample_rate, self.n_mels, self.slice_length, self.width = sample_rate, n_mels, slice_length, width
        self.n_fft, self.hop_length = n_fft, hop_length
        self.device = torch.device(device)
        F = n_fft // 2 + 1
        n = torch.arange(n_fft, dtype=torch.float64)
        ang = 2.0 * math.pi * torch.outer(n, torch.arange(F, dtype=torch.float64)) / n_fft
        self._window = torch.hann_window(n_fft)                                         # periodic hann, as torchaudio builds it
        self._basis = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=1).float().contiguous()   # [n_fft, 2F]
        self._fb = melscale_fbanks_htk(F, 0.0, float(self.sample_rate // 2), self.n_mels, self.sample_rate).float().contiguous()
        self._dev_consts = None

    def _consts(self):
        if self._dev_consts is None:
            self._dev_consts = tuple(t.to(self.device) for t in (self._window, self._basis, self._fb))
        return self._dev_consts

    def load(self, path_to_audio: str, num_chunks: int = 1, startpoint: int = 0, return_wav: bool = False):
        import torchaudio                                   # decoding only
        wav, _ = torchaudio.load(path_to_audio)
        wav = wav.requires_grad_(False)
        if self.slice_length != 0:
            wav = get_slice(wav, self.slice_length, startpoint, num_chunks, self.sample_rate)
        wav = peak_normalizer(wav)
        mel_normed = self.transform_wav(wav)
        return (wav, mel_normed) if return_wav else mel_normed

    def load_batch(self, songlist: List[str], startpoints: List[int] = None) -> torch.Tensor:
        if startpoints is None:
            startpoints = np.zeros(len(songlist))
        samples = [self.load(name, startpoint=startpoint) for name, startpoint in zip(songlist, startpoints)]
        return torch.stack(samples, dim=0).view(-1, 1, self.n_mels, self.width)

    def transform_wav(self, wav: torch.Tensor, return_all: bool = False, clamp: bool = True) -> torch.Tensor:
        """Waveform(s) [.., n_samples] -> log-mel-spectrogram [-1, 1, n_mels, width] on ``self.device``
        (dataloading.py:138-176: frames 1 .. width of log10(mel + 1e-7), clamped at -4)."""
        if return_all:
            raise _L.DRSAError("return_all=True (magnitude / phase / mel as numpy arrays for plotting) is not on this path")
        if not torch.cuda.is_available():
            raise _L.DRSAError("no CUDA device available and no fallback path exists")
        w = wav.detach().to(self.device, torch.float32)
        w = w.reshape(-1, w.shape[-1]).contiguous()
        B, N = w.shape
        n_frames = 1 + N // self.hop_length
        assert n_frames >= self.width + 1, \
            f"width of logmel-spectrogram ({n_frames - 1}) has to equal self.width ({self.width})."
        window, basis, fb = self._consts()
        lib = _L.lib()
        out = torch.empty(B, 1, self.n_mels, self.width, device=self.device)
        with torch.cuda.device(self.device):
            ws = torch.empty(int(_L.check(lib.logmel_transform_workspace_bytes(B, self.n_fft, self.n_mels, self.width))),
                             dtype=torch.uint8, device=self.device)
            _L.check(lib.logmel_transform_wav(w.data_ptr(), window.data_ptr(), basis.data_ptr(), fb.data_ptr(), B, N,
                                              self.n_fft, self.hop_length, self.n_mels, 1, self.width, int(clamp), -4.0,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream), "logmel_transform_wav")
        return out


def shuffle_and_truncate_databatch(data_batch: torch.Tensor, songlist: List[str], N: int, seed: int = 42
                                   ) -> Tuple[torch.Tensor, List[str]]:
    """dataloading.py:179-205: permute with a local torch generator, keep the first N."""
    local_gen = torch.Generator().manual_seed(seed)
    perm = torch.randperm(data_batch.size(0), generator=local_gen)
    data_batch = data_batch[perm.to(data_batch.device)][:N]
    return data_batch, [songlist[i] for i in perm.tolist()][:N]
