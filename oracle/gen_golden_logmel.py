"""Generate tests/golden/logmel_*.npz with the torchaudio objects the reference's Loader builds
(cxai/utils/dataloading.py:63-73) and its transform_wav arithmetic (:151-176).  Run in the build container
(``python -m oracle.gen_golden_logmel``); torchaudio does the work, the inputs are regenerated from seeds."""
import os

import numpy as np
import torch
import torchaudio

from oracle import logmel_ref

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"toy": dict(sample_rate=16000, n_fft=480, hop_length=240, n_mels=64, width=64, seconds=1, B=3, seed=5),
         "gtzan": dict(sample_rate=16000, n_fft=800, hop_length=360, n_mels=128, width=128, seconds=3, B=2, seed=6)}


def main():
    for name, c in CASES.items():
        wav = torch.from_numpy(logmel_ref.synth_wav(c["B"], c["seconds"] * c["sample_rate"], c["seed"], c["sample_rate"]))
        wav2spec = torchaudio.transforms.Spectrogram(n_fft=c["n_fft"], hop_length=c["hop_length"], power=None)
        spec2mel = torchaudio.transforms.MelScale(n_mels=c["n_mels"], n_stft=c["n_fft"] // 2 + 1, sample_rate=c["sample_rate"])
        outs = {}
        for dt in (torch.float32, torch.float64):
            spec = wav2spec(wav.to(dt))
            mel = spec2mel.to(dt)(torch.abs(spec))
            logmel = torch.clamp(torch.log10(mel + 1e-7), -4)
            logmel = logmel[..., 1:c["width"] + 1]
            outs[dt] = logmel.reshape(-1, 1, c["n_mels"], c["width"]).numpy()
        np.savez_compressed(os.path.join(GOLD, f"logmel_{name}.npz"), logmel=outs[torch.float32],
                            logmel_f64=outs[torch.float64].astype(np.float32),
                            wav_checksum=np.array([float(wav.double().sum()), float((wav.double() ** 2).sum())]),
                            **{k: np.array(v) for k, v in c.items()})
        print(name, outs[torch.float32].shape, os.path.getsize(os.path.join(GOLD, f"logmel_{name}.npz")))


if __name__ == "__main__":
    main()
