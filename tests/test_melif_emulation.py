"""CPU checks of the front-end device code: the per-thread phases of
csrc/melif_core.cuh are compiled with g++ (tests/emu/melif_emu.cpp) and run in kernel
order, then compared with the oracle.  Also pins the host-side mel band table."""
import ctypes
import math
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
    so = tmp_path_factory.mktemp("emu") / "melif_emu.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17",
                    f"-I{ROOT / 'interactive_spectrogram_inpainting_b200' / 'csrc'}",
                    "-o", str(so), str(ROOT / "tests" / "emu" / "melif_emu.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    lib.melif_emulate.restype = ctypes.c_int
    return lib


def _run(emu, helper, audio, seg_frames=0, plan=None):
    n_notes, n_samples = audio.shape
    frames = helper.num_frames(n_samples)
    out = np.zeros((n_notes, 2, helper.n_freq, frames), dtype=np.float32)
    a = np.ascontiguousarray(audio.numpy())
    ptr = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    win = helper.window.numpy()
    tw = helper.twiddle.numpy()
    if helper.use_mel_scale:
        ms, mc, mw = helper.mel_start.numpy(), helper.mel_count.numpy(), helper.mel_weight.numpy()
        width = mw.shape[1]
        mel_args = (ptr(ms), ptr(mc), ptr(mw))
    else:
        width, mel_args = 0, (None, None, None)
    rc = emu.melif_emulate(ptr(a), ctypes.c_int64(n_notes), ctypes.c_int64(n_samples),
                           -helper.n_fft if plan == "w32" else helper.n_fft, helper.hop_length, helper.pad_left, frames,
                           1 if helper.drop_bin == "dc" else 0, int(helper.use_mel_scale), width,
                           ctypes.c_float(helper.safelog_eps), ptr(win), ptr(tw), *mel_args,
                           ptr(out), seg_frames,
                           0 if helper.masked_phase_threshold is None else 1,
                           ctypes.c_float(helper.masked_phase_threshold or 0.0),
                           ptr(np.array([v for pair in (helper.output_affine or ((1, 0), (1, 0))) for v in pair],
                                        dtype=np.float32)))
    assert rc == 0
    return torch.from_numpy(out)


def check_against_oracle(got, audio, cfg, max_excluded=0.5):
    """Shared with the GPU parity test.  Tolerance of BASELINE.json's north star: 1e-4
    relative to the channel's max-abs, against the FP64 evaluation of the oracle.

    FP32 cannot meet 1e-4 at *every* position of this transform -- the FP32 torch
    restatement itself misses it at ~0.2 % of positions, where a bin sits > 80 dB under
    its frame's peak (phase = atan2 of rounding noise) or a phase step sits on the +-pi
    wrap (the IF is discontinuous there).  So:
      * log-magnitude: <= 1e-4 x max-abs at >= 99.9 % of positions, never worse than 0.05;
      * IF on well-conditioned positions (FP64 mask: bin within 80 dB of its frame peak,
        step not within 1e-3 rad of the wrap; 60-90 % of positions on the synthetic notes): <= 1e-4 x max-abs
        at >= 99.95 %, never worse than 5e-4;
      * IF everywhere: <= 1e-3 at >= 99.97 %, wrap flips (error ~2) at <= 5e-5.
    Returns the excluded (ill-conditioned) fraction."""
    want = fo.to_spectrogram(audio.double(), cfg)
    stable = fo.stability_mask(audio, cfg, wrap_margin=1e-3, mag_floor=1e-4)
    err0 = (got[:, 0].double() - want[:, 0]).abs()
    tol0 = 1e-4 * want[:, 0].abs().max()
    assert err0.max() < 0.05, err0.max()
    assert (err0 > tol0).double().mean() < 1e-3
    err1 = (got[:, 1].double() - want[:, 1]).abs()
    tol1 = 1e-4 * want[:, 1].abs().max().clamp_min(1.0)
    assert err1[stable].max() <= 5 * tol1, err1[stable].max()
    # (count-based floor: on the small ragged cases 5e-4 is only 2-4 positions)
    assert (err1[stable] > tol1).sum() <= max(8, 5e-4 * err1[stable].numel())
    everywhere = err1.clone()
    if not cfg.use_mel_scale:
        # the purely real bin (DC or Nyquist) has phase exactly 0 or pi: every step sits on
        # the wrap, so its IF sign is decided by the last ulp of atan2 -- excluded here,
        # and flagged by the stability mask above
        everywhere[:, 0 if cfg.drop_bin == "nyquist" else -1] = 0
    assert (everywhere > 10 * tol1).double().mean() < 3e-4
    # (on the small ragged cases 5e-5 is less than one position: allow a handful outright)
    assert (everywhere > 100 * tol1).sum() <= max(4, 5e-5 * everywhere.numel())
    excluded = 1.0 - stable.double().mean().item()
    assert excluded < max_excluded, excluded
    return excluded


def test_band_table_equals_dense_recipe():
    cfg = fo.FrontEndConfig()
    starts, counts, weights = sh.mel_band_table(2048, 16000, 0.0, 8000.0, 700.0, 1.5)
    np.testing.assert_allclose(sh.dense_mel_matrix(starts, counts, weights),
                               fo.linear_to_mel_matrix(cfg), rtol=0, atol=1e-12)
    assert weights.shape[1] <= 8


@pytest.mark.parametrize("use_mel", [True, False])
def test_emulated_kernel_matches_oracle_2048(emu, use_mel):
    audio = synthetic.synthetic_notes(2)
    helper = (sh.MelSpectrogramsHelper() if use_mel else sh.SpectrogramsHelper())
    got = _run(emu, helper, audio)
    assert got.shape == (2, 2, 1024, 128)
    check_against_oracle(got, audio, fo.FrontEndConfig(use_mel_scale=use_mel))


@pytest.mark.parametrize("use_mel", [True, False])
def test_emulated_one_warp_plan_matches_oracle_and_the_generic_plan(emu, use_mel):
    """The warp-specialised kernel's transform (PlanW32: one warp per frame pair, two radix-32
    passes with one exchange through shared memory): same oracle bars, segments and ragged
    lengths included, and the same spectrum as the 16 x 16 x 4 plan up to FP32 rounding."""
    audio = synthetic.synthetic_notes(2)
    helper = (sh.MelSpectrogramsHelper() if use_mel else sh.SpectrogramsHelper())
    got = _run(emu, helper, audio, plan="w32")
    assert got.shape == (2, 2, 1024, 128)
    check_against_oracle(got, audio, fo.FrontEndConfig(use_mel_scale=use_mel))
    generic = _run(emu, helper, audio)
    d0 = (got[:, 0] - generic[:, 0]).abs()
    tol = 1e-4 * generic[:, 0].abs().max()                              # the oracle bar's scale
    assert (d0 > tol).double().mean() < 2e-3 and d0.max() < 0.05, ((d0 > tol).double().mean(), d0.max())
    for seg in (16, 40):
        assert torch.equal(_run(emu, helper, audio[:1], seg_frames=seg, plan="w32"), got[:1])
    short = synthetic.synthetic_notes(1, n_samples=9001)
    check_against_oracle(_run(emu, helper, short, plan="w32"), short, fo.FrontEndConfig(use_mel_scale=use_mel))


@pytest.mark.parametrize("n_fft,hop,samples", [(1024, 256, 9000), (512, 128, 4099)])
def test_emulated_kernel_other_sizes_and_ragged_lengths(emu, n_fft, hop, samples):
    audio = synthetic.synthetic_notes(1, n_samples=samples)
    helper = sh.MelSpectrogramsHelper(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    got = _run(emu, helper, audio)
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    assert got.shape[-1] == fo.frame_geometry(cfg, samples)[2]
    check_against_oracle(got, audio, cfg)


def test_emulated_segments_equal_whole_note(emu):
    """A note split into frame segments (each re-deriving the previous frame's spectrum with
    a look-back transform) gives bit-identical output to the unsplit note."""
    audio = synthetic.synthetic_notes(1)
    helper = sh.MelSpectrogramsHelper()
    whole = _run(emu, helper, audio)
    for seg in (16, 44, 64):
        assert torch.equal(_run(emu, helper, audio, seg_frames=seg), whole)


def test_emulated_fused_epilogue_matches_oracle(emu):
    audio = synthetic.synthetic_notes(1)
    kw = dict(masked_phase_threshold=-3.0, output_affine=((0.1, 0.5), (2.0, -0.25)))
    got = _run(emu, sh.MelSpectrogramsHelper(**kw), audio)
    plain = _run(emu, sh.MelSpectrogramsHelper(), audio)
    want = fo.epilogue(plain, **kw)
    assert torch.allclose(got, want, rtol=0, atol=1e-6)
    assert (got[:, 1][plain[:, 0] < -3.0] == -0.25).all()      # masked phase -> bias only


def test_emulated_kernel_nyquist_knob(emu):
    audio = synthetic.synthetic_notes(1, n_samples=16000)
    helper = sh.SpectrogramsHelper(drop_bin="nyquist", window_periodic=False)
    got = _run(emu, helper, audio)
    cfg = fo.FrontEndConfig(use_mel_scale=False, drop_bin="nyquist", window_periodic=False)
    check_against_oracle(got, audio, cfg)


def test_emulated_mel_mode_with_dropped_nyquist(emu):
    """Item 0 of the polar step owns bin M/2 and the real bin that is kept (DC here)."""
    audio = synthetic.synthetic_notes(1, n_samples=16000)
    helper = sh.MelSpectrogramsHelper(drop_bin="nyquist")
    check_against_oracle(_run(emu, helper, audio), audio, fo.FrontEndConfig(drop_bin="nyquist"))


@pytest.mark.parametrize("n_fft,hop,samples", [(512, 125, 3001), (1024, 250, 5000), (2048, 500, 40000)])
def test_emulated_kernel_odd_and_unaligned_hops(emu, n_fft, hop, samples):
    audio = synthetic.synthetic_notes(1, n_samples=samples)
    helper = sh.MelSpectrogramsHelper(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    check_against_oracle(_run(emu, helper, audio), audio, cfg)
