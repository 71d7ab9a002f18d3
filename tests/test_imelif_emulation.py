"""CPU checks of the inverse front end's device code: the per-thread phases of
csrc/imelif_core.cuh are compiled with g++ (tests/emu/imelif_emu.cpp) and run in kernel
order, then compared with the oracle's ``to_audio``.  Also pins the host-side tables."""
import ctypes
import pathlib
import subprocess

import numpy as np
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import spectrograms_helper as sh
from interactive_spectrogram_inpainting_b200.utils import synthetic
from oracle import frontend_oracle as fo

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "imelif_emu.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17",
                    f"-I{ROOT / 'interactive_spectrogram_inpainting_b200' / 'csrc'}",
                    "-o", str(so), str(ROOT / "tests" / "emu" / "imelif_emu.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    lib.imelif_emulate.restype = ctypes.c_int
    return lib


def _run(emu, helper, spec, seg_frames=0, vec_out=1):
    n_notes, _, n_freq, frames = spec.shape
    n_samples = helper.hop_length * frames - helper.pad_left
    out = np.full((n_notes, n_samples), np.nan, dtype=np.float32)
    x = np.ascontiguousarray(spec.numpy(), dtype=np.float32)
    ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    win, tw = helper.window.numpy(), helper.twiddle.numpy()
    ola = helper._ola_scale(frames, "cpu").numpy()
    if helper.use_mel_scale:
        bs, bc, bw = helper.inv_start.numpy(), helper.inv_count.numpy(), helper.inv_weight.numpy()
        width, band = bw.shape[1], (ptr(bs), ptr(bc), ptr(bw))
    else:
        width, band = 0, (None, None, None)
    affine = np.array([v for pair in (helper.input_affine or ((1, 0), (1, 0))) for v in pair], dtype=np.float32)
    rc = emu.imelif_emulate(ptr(x), ctypes.c_int64(n_notes), helper.n_fft, helper.hop_length,
                            helper.pad_left, frames, 1 if helper.drop_bin == "dc" else 0,
                            int(helper.use_mel_scale), width, ctypes.c_float(helper.safelog_eps),
                            ptr(win), ptr(tw), *band, ptr(ola), ptr(affine), ptr(out),
                            ctypes.c_int64(n_samples), seg_frames, vec_out)
    assert rc == 0
    return torch.from_numpy(out)


def check_audio_against_oracle(got, spec, cfg, input_affine=None, tol=1e-4):
    """Shared with the GPU parity test.  BASELINE.json's tolerance: 1e-4 relative to the
    signal's max-abs, against the FP64 evaluation of the oracle -- at every sample."""
    want = fo.to_audio(spec.double(), cfg, input_affine=input_affine)
    assert got.shape == want.shape
    assert torch.isfinite(got).all()
    scale = want.abs().max().clamp_min(1e-12)
    err = (got.double() - want).abs().max() / scale
    assert err <= tol, float(err)
    return float(err)


def _spec_of_notes(n, cfg, seed=0):
    audio = synthetic.synthetic_notes(n, seed=seed) if "seed" in synthetic.synthetic_notes.__code__.co_varnames \
        else synthetic.synthetic_notes(n)
    return fo.to_spectrogram(audio.double(), cfg).float()


def _random_spec(n, n_freq, frames, seed):
    g = torch.Generator().manual_seed(seed)
    logmag = torch.randn(n, n_freq, frames, generator=g) * 2.0 - 3.0
    ifreq = torch.rand(n, n_freq, frames, generator=g) * 2.0 - 1.0
    return torch.stack([logmag, ifreq], 1)


@pytest.mark.parametrize("mel", [True, False])
def test_emulated_inverse_matches_oracle_on_notes(emu, mel):
    cfg = fo.FrontEndConfig(use_mel_scale=mel)
    helper = (sh.MelSpectrogramsHelper if mel else sh.SpectrogramsHelper)()
    spec = _spec_of_notes(2, cfg)
    got = _run(emu, helper, spec)
    check_audio_against_oracle(got, spec, cfg)


@pytest.mark.parametrize("mel", [True, False])
@pytest.mark.parametrize("seg_frames", [8, 16, 48])
def test_segments_reproduce_the_whole_note(emu, mel, seg_frames):
    """A CTA that starts mid-note seeds its phases from the FP64 prefix sums and
    re-synthesises the frames that reach into its first hop."""
    cfg = fo.FrontEndConfig(use_mel_scale=mel)
    helper = (sh.MelSpectrogramsHelper if mel else sh.SpectrogramsHelper)()
    spec = _random_spec(1, 1024, 128, seed=3)
    got = _run(emu, helper, spec, seg_frames=seg_frames)
    check_audio_against_oracle(got, spec, cfg)
    whole = _run(emu, helper, spec)
    assert (got - whole).abs().max() <= 2e-5 * whole.abs().max()


@pytest.mark.parametrize("n_fft,hop,frames,vec", [(1024, 256, 36, 1), (512, 128, 30, 1), (2048, 510, 21, 0),
                                                   (1024, 256, 17, 0), (512, 512, 9, 1), (2048, 512, 6, 1)])
@pytest.mark.parametrize("mel", [True, False])
def test_other_geometries(emu, n_fft, hop, frames, vec, mel):
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft, use_mel_scale=mel)
    cls = sh.MelSpectrogramsHelper if mel else sh.SpectrogramsHelper
    helper = cls(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    spec = _random_spec(2, n_fft // 2, frames, seed=n_fft + frames)
    for seg in (0, 8):
        got = _run(emu, helper, spec, seg_frames=seg, vec_out=vec)
        check_audio_against_oracle(got, spec, cfg)


def test_knobs_nyquist_bin_padding_and_affine(emu):
    cfg = fo.FrontEndConfig(use_mel_scale=True, drop_bin="nyquist", pad_left=1024)
    helper = sh.MelSpectrogramsHelper(drop_bin="nyquist", pad_left=1024)
    helper.input_affine = ((2.0, -1.0), (0.5, 0.1))
    spec = _random_spec(1, 1024, 24, seed=11)
    for vec in (0, 1):
        got = _run(emu, helper, spec, seg_frames=8, vec_out=vec)
        check_audio_against_oracle(got, spec, cfg, input_affine=helper.input_affine)


def test_inverse_band_table_is_the_normalised_transpose():
    helper = sh.MelSpectrogramsHelper()
    n = helper.n_freq
    dense = np.zeros((n, n))
    for l in range(n):
        s, c = int(helper.inv_start[l]), int(helper.inv_count[l])
        dense[s:s + c, l] = helper.inv_weight[l, :c].double().numpy()
    want = fo.mel_to_linear_matrix(fo.FrontEndConfig())
    assert np.abs(dense - want).max() < 1e-6
    assert int(helper.inv_count.max()) <= 8


def test_round_trip_of_the_linear_transform(emu):
    """forward (oracle) -> inverse (device code) returns the audio up to the eps in log(|X|+eps)
    and the dropped DC bin."""
    cfg = fo.FrontEndConfig(use_mel_scale=False)
    audio = synthetic.synthetic_notes(1)
    spec = fo.to_spectrogram(audio.double(), cfg).float()
    back = _run(emu, sh.SpectrogramsHelper(), spec)
    assert (back - audio).abs().max() < 5e-3
