"""Pins oracle/logmel_ref.py (numpy fp64 restatement of Loader.transform_wav) against golden vectors produced by the
torchaudio transforms the reference builds (oracle/gen_golden_logmel.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import logmel_ref


@pytest.mark.parametrize("name", ["toy", "gtzan"])
def test_logmel_oracle_matches_torchaudio(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"logmel_{name}.npz"))
    c = {k: int(g[k]) for k in ("sample_rate", "n_fft", "hop_length", "n_mels", "width", "seconds", "B", "seed")}
    wav = logmel_ref.synth_wav(c["B"], c["seconds"] * c["sample_rate"], c["seed"], c["sample_rate"])
    np.testing.assert_allclose([wav.astype(np.float64).sum(), (wav.astype(np.float64) ** 2).sum()], g["wav_checksum"], rtol=1e-9)
    out = logmel_ref.transform_wav(wav, c["sample_rate"], c["n_fft"], c["hop_length"], c["n_mels"], c["width"])
    assert out.shape == g["logmel"].shape == (c["B"], 1, c["n_mels"], c["width"])
    # torchaudio in fp64 (stored as fp32): agreement to rounding of the stored values
    np.testing.assert_allclose(out, g["logmel_f64"], atol=2e-6, rtol=1e-6)
    # torchaudio in fp32 (what the reference runs): its own FFT rounding shows up in the quiet bins
    assert np.abs(out - g["logmel"]).max() < 5e-3
    assert np.mean(np.abs(out - g["logmel"])) < 2e-5
    assert out.min() >= -4.0 and (out == -4.0).any() == (g["logmel"] == -4.0).any()
