"""The front-end restatement (PARITY UNPINNED: GANsynth_pytorch is not in the
reference tree) against its own regression fixture and closed-form properties."""
import math

import numpy as np
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from oracle import frontend_oracle as fo


def test_shapes_match_the_reference_call_sites():
    cfg = fo.FrontEndConfig()
    assert fo.frame_geometry(cfg, 64000) == (1536, 1536, 128)   # Inference.ipynb:71
    spec = fo.to_spectrogram(synthetic.synthetic_notes(1), cfg)
    assert spec.shape == (1, 2, 1024, 128) and spec.dtype == torch.float32


def test_regression_fixture(golden_dir):
    g = np.load(golden_dir / "frontend_unpinned.npz")
    audio = synthetic.synthetic_notes(2)
    np.testing.assert_array_equal(audio[:, :4096].numpy(), g["audio_head"])
    mel = fo.to_spectrogram(audio, fo.FrontEndConfig())[:, :, ::8, ::4].numpy()
    lin = fo.to_spectrogram(audio, fo.FrontEndConfig(use_mel_scale=False))[:, :, ::8, ::4].numpy()
    for got, want in ((mel, g["mel"]), (lin, g["lin"])):
        assert np.abs(got[:, 0] - want[:, 0]).max() < 2e-3
        assert np.mean(np.abs(got[:, 1] - want[:, 1]) > 1e-3) < 1e-3


def test_pure_tone_if_is_its_frequency():
    """A stationary sinusoid at f advances hop*2*pi*f/fs per frame: the linear IF
    in the peak bin is that step wrapped to (-pi, pi] over pi."""
    cfg = fo.FrontEndConfig(use_mel_scale=False)
    f = 1000.0
    t = torch.arange(64000, dtype=torch.float64) / cfg.fs_hz
    spec = fo.to_spectrogram(torch.sin(2 * math.pi * f * t)[None], cfg)
    k = int(round(f / (cfg.fs_hz / cfg.n_fft))) - 1            # DC dropped
    step = (cfg.hop_length * 2 * math.pi * f / cfg.fs_hz + math.pi) % (2 * math.pi) - math.pi
    mid = spec[0, 1, k, 8:120].double()
    assert (mid - step / math.pi).abs().max() < 1e-3
    assert abs(spec[0, 0, k, 64].item() - math.log(0.5 * 1024 + 1e-6)) < 1e-2


def test_mel_matrix_is_banded_and_nonnegative():
    m = fo.linear_to_mel_matrix(fo.FrontEndConfig())
    assert m.shape == (1024, 1024) and (m >= 0).all() and (m[0] == 0).all()
    for j in range(1024):
        nz = np.nonzero(m[:, j])[0]
        if len(nz):
            assert nz[-1] - nz[0] + 1 == len(nz) <= 8          # contiguous band


def test_unwrap_matches_numpy():
    g = torch.Generator().manual_seed(3)
    ph = (torch.rand(5, 40, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    np.testing.assert_allclose(fo.unwrap_time(ph).numpy(), np.unwrap(ph.numpy(), axis=-1),
                               atol=1e-12)
