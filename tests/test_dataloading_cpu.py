"""Host-side pieces of the audio frontend (no GPU): ``get_slice`` against the reference's slicing (utils/sound.py:8-44)."""
import math

import numpy as np
import pytest
import torch

from cxai.utils.dataloading import get_slice, peak_normalizer


def _ref_get_slice(wav, slice_length=6, start_point=0, num_chunks=1, sample_rate=16000):
    """utils/sound.py:31-44 restated (test infrastructure)."""
    window_size = int(slice_length * sample_rate)
    if num_chunks > 1:
        hop = int(math.floor(((29 - slice_length) / (num_chunks - 1)) * 10 ** 1) / 10 ** 1 * sample_rate)
        out = wav[:, :29 * sample_rate].unfold(1, window_size, hop).reshape(-1, 1, window_size)
        assert out.shape[0] == num_chunks
        return out
    s0 = int(start_point * sample_rate)
    return wav[:, s0:s0 + window_size]


@pytest.mark.parametrize("slice_length,num_chunks", [(3, 10), (6, 3), (6, 5), (3, 2)])
def test_get_slice_overlapping_chunks_match_reference(slice_length, num_chunks):
    sr = 1000
    wav = torch.randn(1, int(29.3 * sr), generator=torch.Generator().manual_seed(0))      # a ~29.3 s clip, like GTZAN
    got = get_slice(wav, slice_length, 7, num_chunks, sr)                                    # start_point is ignored here
    want = _ref_get_slice(wav, slice_length, 7, num_chunks, sr)
    assert got.shape == (num_chunks, 1, slice_length * sr)
    np.testing.assert_array_equal(got.numpy(), want.numpy())
    # the windows stay inside the first 29 s and overlap or tile with the rounded-down hop
    hop = int(math.floor((29 - slice_length) / (num_chunks - 1) * 10) / 10 * sr)
    np.testing.assert_array_equal(got[1, 0, :5].numpy(), wav[0, hop:hop + 5].numpy())


def test_get_slice_single_window_and_range_check():
    sr = 100
    wav = torch.arange(2 * 30 * sr, dtype=torch.float32).reshape(2, -1)                    # two channels stay separate
    got = get_slice(wav, 6, 4, 1, sr)
    assert got.shape == (2, 6 * sr)
    np.testing.assert_array_equal(got.numpy(), wav[:, 4 * sr:10 * sr].numpy())
    with pytest.raises(AssertionError):
        get_slice(wav, 6, 30 * sr, 1, sr)
    np.testing.assert_allclose(peak_normalizer(got).abs().amax(dim=-1).numpy(), 1.0)


def test_get_slice_and_best_run_match_reference_fixture(golden_dir, tmp_path):
    """The same two host utilities against the outputs of the reference's OWN functions (utils/sound.py:8-44,
    utils/evaluation.py:107-141; oracle/gen_golden_lrp.py `hostutils`)."""
    import os
    from cxai.utils.evaluation import get_best_run
    g = np.load(os.path.join(golden_dir, "lrp_hostutils.npz"))
    sr = int(g["sr"])
    wav = torch.randn(1, int(g["wav_len"]), generator=torch.Generator().manual_seed(int(g["wav_seed"])))
    for sl, nc in g["slice_combos"].tolist():
        np.testing.assert_array_equal(get_slice(wav, sl, 7, nc, sr).numpy(), g[f"slices_{sl}_{nc}"])
    np.testing.assert_array_equal(get_slice(wav, 6, 4, 1, sr).numpy(), g["slice_single"])
    for r, ls in zip((1, 2, 3), g["tree_losses"].tolist()):
        os.makedirs(tmp_path / f"run{r}")
        with open(tmp_path / f"run{r}" / "train_stats.csv", "w") as f:
            f.write(",loss\n" + "".join(f"{i},{v}\n" for i, v in enumerate(ls)))
    run, loss, crel, path, losses = get_best_run(str(tmp_path))
    assert run == int(g["best_run"]) and loss == float(g["best_loss"]) and os.path.basename(path) == str(g["best_dir"])
    assert len(crel) == int(g["n_concept_relevances"]) and len(losses) == 3
