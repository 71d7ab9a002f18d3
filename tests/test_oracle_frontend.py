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


def test_inverse_regression_fixture(golden_dir):
    """``to_audio`` restatement (unpinned like the forward transform) against its fixture."""
    g = np.load(golden_dir / "inverse_unpinned.npz")
    gen = torch.Generator().manual_seed(20200117)
    spec = torch.stack([torch.randn(2, 1024, 24, generator=gen) * 2.0 - 3.0,
                        torch.rand(2, 1024, 24, generator=gen) * 2.0 - 1.0], 1)
    np.testing.assert_array_equal(spec[:, :, :8, :8].numpy(), g["spec_corner"])
    for mel, key in ((True, "mel"), (False, "lin")):
        got = fo.to_audio(spec.double(), fo.FrontEndConfig(use_mel_scale=mel))[:, ::7].numpy()
        assert got.shape == g[key].shape == (2, 1536)
        assert np.abs(got - g[key]).max() <= 1e-5 * np.abs(g[key]).max()


def test_inverse_of_forward_is_the_audio():
    """Linear scale: forward then inverse returns the note up to the eps of log(|X| + eps) and
    the dropped DC bin; the padded frames make every kept sample fully overlapped."""
    cfg = fo.FrontEndConfig(use_mel_scale=False)
    audio = synthetic.synthetic_notes(1).double()
    back = fo.to_audio(fo.to_spectrogram(audio, cfg), cfg)
    assert back.shape == audio.shape
    assert (back - audio).abs().max() < 5e-3


def test_mel_to_linear_matrix_undoes_the_filterbank_on_smooth_spectra():
    cfg = fo.FrontEndConfig()
    fwd, back = fo.linear_to_mel_matrix(cfg), fo.mel_to_linear_matrix(cfg)
    assert back.shape == (1024, 1024) and (back >= 0).all()
    flat = np.ones(1024)
    flat[0] = 0.0                                              # the zeroed first linear row
    rebuilt = (flat @ fwd) @ back                              # linear -> mel -> linear
    # the column normalisation makes a flat spectrum come back exactly; the top bin sits on the
    # upper band edge (8000 Hz = Nyquist) and receives no weight at all
    assert np.abs(rebuilt[1:-1] - 1.0).max() < 1e-12 and rebuilt[-1] == 0.0 and rebuilt[0] == 0.0


def test_stft_agrees_with_scipy():
    """The restatement's framing / window / transform against an independent implementation
    (scipy.signal.stft, no boundary extension, periodic Hann), FP64."""
    import scipy.signal
    cfg = fo.FrontEndConfig(use_mel_scale=False)
    audio = synthetic.synthetic_notes(1, n_samples=16000).double()
    ours = fo.stft(audio, cfg)[0].numpy()                                  # bins 1..1024
    pad_l, pad_r, frames = fo.frame_geometry(cfg, 16000)
    padded = np.pad(audio[0].numpy(), (pad_l, pad_r))
    win = scipy.signal.get_window("hann", cfg.n_fft, fftbins=True)
    _, _, z = scipy.signal.stft(padded, window=win, nperseg=cfg.n_fft, noverlap=cfg.n_fft - cfg.hop_length,
                                boundary=None, padded=False, return_onesided=True)
    z = z * win.sum()                                                      # scipy normalises by the window sum
    assert z.shape == (cfg.n_fft // 2 + 1, frames)
    assert np.abs(ours - z[1:]).max() <= 1e-9 * np.abs(z).max()


def test_mel_matrix_agrees_with_a_direct_evaluation_of_the_published_formula():
    """magenta ``linear_to_mel_weight_matrix`` written out bin by bin in plain Python for a few
    columns: triangle between equally spaced mel edges, HTK-style mel with break frequency 700 Hz
    and Q = 1127, narrow triangles widened to 1.5 linear bins (arcsinh re-centring)."""
    cfg = fo.FrontEndConfig()
    m = fo.linear_to_mel_matrix(cfg)
    n, nyq, brk, q = 1024, 8000.0, 700.0, 1127.0
    mel = lambda f: q * math.log(1.0 + f / brk)
    hz = lambda v: brk * (math.exp(v / q) - 1.0)
    edges = [mel(0.0) + (mel(8000.0) - mel(0.0)) * i / (n + 1) for i in range(n + 2)]
    for j in (0, 3, 40, 300, 700, 1023):
        lo, mid, hi = edges[j], edges[j + 1], edges[j + 2]
        floor = 1.5 * nyq / n
        if hz(hi) - hz(lo) < floor:
            half = q * math.asinh(0.5 * floor / (hz(mid) + brk))
            lo, hi = mid - half, mid + half
        for k in range(1, n):
            f = nyq * k / (n - 1)
            want = max(0.0, min((f - hz(lo)) / (hz(mid) - hz(lo)), (hz(hi) - f) / (hz(hi) - hz(mid))))
            assert abs(m[k, j] - want) < 1e-12, (k, j)
        assert m[0, j] == 0.0
