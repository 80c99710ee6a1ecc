"""GPU parity of the log-mel frontend (cxai.utils.dataloading.Loader.transform_wav on libdrsa_b200.so) against the
oracle and the torchaudio golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import logmel_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["toy", "gtzan"])
def test_transform_wav_matches_golden_and_oracle(golden_dir, name):
    from cxai.utils.dataloading import Loader
    g = np.load(os.path.join(golden_dir, f"logmel_{name}.npz"))
    c = {k: int(g[k]) for k in ("sample_rate", "n_fft", "hop_length", "n_mels", "width", "seconds", "B", "seed")}
    wav = logmel_ref.synth_wav(c["B"], c["seconds"] * c["sample_rate"], c["seed"], c["sample_rate"])
    loader = Loader(case=name)
    assert (loader.n_fft, loader.hop_length, loader.n_mels, loader.width) == (c["n_fft"], c["hop_length"], c["n_mels"], c["width"])
    out = loader.transform_wav(torch.from_numpy(wav)).cpu().numpy()
    want, mel = logmel_ref.transform_wav(wav, c["sample_rate"], c["n_fft"], c["hop_length"], c["n_mels"], c["width"],
                                         return_mel=True)
    assert out.shape == want.shape
    # log10 amplifies the fp32 rounding of the DFT in bins 60 dB and more below the peak: compare where the mel energy
    # carries information, bound everything else
    loud = mel > 1e-4 * mel.max()
    assert np.abs(out - want)[loud].max() < 2e-4
    assert np.abs(out - want).max() < 1e-2
    assert np.abs(out - g["logmel"]).max() < 1e-2 and np.mean(np.abs(out - g["logmel"])) < 5e-5      # torchaudio fp32
    assert out.min() >= -4.0
    # unclamped variant and a single unbatched waveform
    raw = loader.transform_wav(torch.from_numpy(wav[0]), clamp=False).cpu().numpy()
    want_raw = logmel_ref.transform_wav(wav[:1], c["sample_rate"], c["n_fft"], c["hop_length"], c["n_mels"], c["width"], clamp=False)
    assert raw.shape == (1, 1, c["n_mels"], c["width"])
    assert np.abs(raw - want_raw)[loud[:1]].max() < 2e-4


def test_logmel_feeds_the_cnn():
    """waveform -> log-mel (device) -> LRP context extraction: the staged spectrogram is what get_intermediate reads."""
    from cxai.utils.dataloading import Loader
    from cxai.utils.constants import LRP_NAME_MAP_TOY
    from cxai.xai.explain.rules import NameMapComposite
    from cxai.xai.drsa.preprocessing import get_intermediate
    from oracle import lrp_ref
    wav = logmel_ref.synth_wav(4, 16000, 9)
    x = Loader(case="toy").transform_wav(torch.from_numpy(wav))
    assert x.shape == (4, 1, 64, 64) and x.is_cuda
    net = lrp_ref.toy_model(seed=0, last=64)
    a, R = get_intermediate(net, x, NameMapComposite(LRP_NAME_MAP_TOY), net.features[13], 1)
    aw, Rw = lrp_ref.get_intermediate(net, x.cpu(), LRP_NAME_MAP_TOY, net.features[13], 1)
    rel = lambda p, q: float(((p.double().cpu() - q).flatten(1).norm(dim=1) / q.flatten(1).norm(dim=1).clamp(min=1e-30)).max())
    assert rel(a, aw) < 1e-5 and rel(R, Rw) < 1e-4
