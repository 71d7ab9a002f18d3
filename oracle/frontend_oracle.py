"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the GANSynth-style front end.

PARITY UNPINNED.  The arithmetic of ``SpectrogramsHelper.to_spectrogram`` /
``MelSpectrogramsHelper.to_spectrogram`` does not live in the reference tree: the
reference only instantiates those classes from the third-party package
``GANsynth_pytorch`` (``interactive_spectrogram_inpainting/utils/misc.py:5-29``,
``train_vqvae.py:61-79``), which is un-vendored, un-pinned (no requirements /
lock file / submodule; README.md:9-14 pins only PyTorch>=1.6, torchaudio>=0.6)
and not installed here.  The reference holds no test or golden vector at this
boundary.  This file therefore restates the *published* GANSynth recipe
(Engel et al. 2019; magenta ``specgrams_helper.py`` / ``spectral_ops.py``) with
the parameters the reference does pin at its call sites:

* fs 16 kHz, n_fft = window 2048, hop 512          (train_vqvae.py:56-58,457-461)
* output ``[B, 2, 1024 freq, 128 time]`` for 64 000 samples
                                          (Inference.ipynb:71, flask_server.py:891-896)
* mel range 0-8000 Hz, break frequency 700 Hz, bin-width threshold factor 1.5
                                                        (train_vqvae.py:474-481)
* ``safelog_eps`` attribute                               (train_vqvae.py:711)

Every choice that cannot be verified against the missing source is a field of
``FrontEndConfig`` (padding, dropped bin, window periodicity).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
import this file.
"""
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

_MEL_HIGH_FREQUENCY_Q = 1127.0


@dataclass(frozen=True)
class FrontEndConfig:
    fs_hz: int = 16000
    n_fft: int = 2048
    hop_length: int = 512
    window_length: int = 2048
    use_mel_scale: bool = True
    lower_edge_hertz: float = 0.0
    upper_edge_hertz: float = 8000.0
    mel_break_frequency_hertz: float = 700.0
    mel_bin_width_threshold_factor: float = 1.5
    safelog_eps: float = 1e-6
    # ---- knobs for choices the missing source would decide ----
    pad_left: Optional[int] = None      # default n_fft - hop (GANSynth _get_padding)
    n_frames: Optional[int] = None      # default ceil((T + pad_left) / hop) -> 128 @ 64000
    drop_bin: str = "dc"                # "dc" (GANSynth discard_dc=True) or "nyquist"
    window_periodic: bool = True        # tf.signal.hann_window default

    @property
    def n_freq(self) -> int:
        return self.n_fft // 2


def frame_geometry(cfg: FrontEndConfig, n_samples: int) -> Tuple[int, int, int]:
    """(pad_left, pad_right, n_frames).  GANSynth pads so that the inverse STFT
    length hop*(frames-1)+n_fft covers the audio; left pad is n_fft-hop."""
    pad_l = cfg.n_fft - cfg.hop_length if cfg.pad_left is None else cfg.pad_left
    if cfg.n_frames is None:
        frames = max(1, math.ceil((n_samples + pad_l) / cfg.hop_length))
    else:
        frames = cfg.n_frames
    total = cfg.hop_length * (frames - 1) + cfg.n_fft
    pad_r = total - n_samples - pad_l
    if pad_r < 0:
        raise ValueError("n_frames too small for the audio length")
    return pad_l, pad_r, frames


def analysis_window(cfg: FrontEndConfig, dtype=torch.float32) -> torch.Tensor:
    w = torch.hann_window(cfg.window_length, periodic=cfg.window_periodic,
                          dtype=torch.float64)
    if cfg.window_length < cfg.n_fft:       # centred zero-padding like torch.stft
        left = (cfg.n_fft - cfg.window_length) // 2
        w = torch.nn.functional.pad(w, (left, cfg.n_fft - cfg.window_length - left))
    return w.to(dtype)


def hertz_to_mel(f, brk):
    return _MEL_HIGH_FREQUENCY_Q * np.log1p(np.asarray(f, dtype=np.float64) / brk)


def mel_to_hertz(m, brk):
    return brk * np.expm1(np.asarray(m, dtype=np.float64) / _MEL_HIGH_FREQUENCY_Q)


def linear_to_mel_matrix(cfg: FrontEndConfig) -> np.ndarray:
    """float64 ``[n_freq linear, n_freq mel]`` triangular filterbank (magenta
    ``spectral_ops.linear_to_mel_weight_matrix`` as GANSynth calls it: as many
    mel bins as linear bins, first linear row zeroed, narrow low-frequency
    triangles widened to ``threshold_factor`` x the linear resolution)."""
    n_bins = cfg.n_freq
    n_mel = cfg.n_freq
    brk = cfg.mel_break_frequency_hertz
    nyquist = cfg.fs_hz / 2.0
    lin_hz = np.linspace(0.0, nyquist, n_bins)[1:, None]
    edges = np.linspace(hertz_to_mel(cfg.lower_edge_hertz, brk),
                        hertz_to_mel(cfg.upper_edge_hertz, brk), n_mel + 2)
    lo_mel, mid_mel, hi_mel = edges[:-2].copy(), edges[1:-1].copy(), edges[2:].copy()
    width_floor = cfg.mel_bin_width_threshold_factor * nyquist / float(n_bins)
    for j in range(n_mel):
        mid_hz = mel_to_hertz(mid_mel[j], brk)
        if mel_to_hertz(hi_mel[j], brk) - mel_to_hertz(lo_mel[j], brk) < width_floor:
            r = 0.5 * width_floor / (mid_hz + brk)
            half = _MEL_HIGH_FREQUENCY_Q * np.log(r + np.sqrt(1.0 + r * r))
            lo_mel[j] = mid_mel[j] - half
            hi_mel[j] = mid_mel[j] + half
    lo_hz = mel_to_hertz(lo_mel, brk)[None, :]
    mid_hz = mel_to_hertz(mid_mel, brk)[None, :]
    hi_hz = mel_to_hertz(hi_mel, brk)[None, :]
    rising = (lin_hz - lo_hz) / (mid_hz - lo_hz)
    falling = (hi_hz - lin_hz) / (hi_hz - mid_hz)
    tri = np.maximum(0.0, np.minimum(rising, falling))
    return np.pad(tri, [[1, 0], [0, 0]])


def _diff_time(x: torch.Tensor) -> torch.Tensor:
    return x[..., 1:] - x[..., :-1]


def unwrap_time(phase: torch.Tensor) -> torch.Tensor:
    """numpy-style unwrap along the last (time) axis (magenta ``spectral_ops.unwrap``)."""
    dd = _diff_time(phase)
    two_pi = 2.0 * math.pi
    ddmod = torch.remainder(dd + math.pi, two_pi) - math.pi
    ddmod = torch.where((ddmod == -math.pi) & (dd > 0),
                        torch.full_like(ddmod, math.pi), ddmod)
    correction = ddmod - dd
    correction = torch.where(dd.abs() < math.pi, torch.zeros_like(correction), correction)
    csum = torch.cumsum(correction, dim=-1)
    csum = torch.cat([torch.zeros_like(phase[..., :1]), csum], dim=-1)
    return phase + csum


def instantaneous_frequency(phase: torch.Tensor) -> torch.Tensor:
    """Finite difference of the time-unwrapped phase, first frame kept, over pi."""
    unwrapped = unwrap_time(phase)
    d = _diff_time(unwrapped)
    return torch.cat([unwrapped[..., :1], d], dim=-1) / math.pi


def stft(audio: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """complex ``[B, n_freq, frames]`` (one bin dropped per ``cfg.drop_bin``)."""
    pad_l, pad_r, frames = frame_geometry(cfg, audio.shape[-1])
    padded = torch.nn.functional.pad(audio, (pad_l, pad_r))
    spec = torch.stft(padded, cfg.n_fft, hop_length=cfg.hop_length,
                      win_length=cfg.n_fft, window=analysis_window(cfg, audio.dtype),
                      center=False, return_complex=True)
    assert spec.shape[-1] == frames
    return spec[:, 1:, :] if cfg.drop_bin == "dc" else spec[:, :-1, :]


def to_linear_spectrogram(audio: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """``[B, 2, n_freq, frames]``: log(|X| + eps) and IF of angle(X)."""
    s = stft(audio, cfg)
    logmag = torch.log(s.abs() + cfg.safelog_eps)
    ifreq = instantaneous_frequency(torch.angle(s))
    return torch.stack([logmag, ifreq], dim=1)


def linear_to_mel(spec: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """magenta ``specgrams_to_melspecgrams``: squared magnitude and the
    re-integrated phase are both projected with the same filterbank; the mel IF
    is the (re-unwrapped) time difference of the projected phase."""
    logmag, ifreq = spec[:, 0], spec[:, 1]
    bank = torch.from_numpy(linear_to_mel_matrix(cfg)).to(spec.dtype)   # [lin, mel]
    mag2 = torch.exp(2.0 * logmag)
    phase = torch.cumsum(ifreq * math.pi, dim=-1)
    mel_mag2 = torch.matmul(bank.t(), mag2)
    mel_phase = torch.matmul(bank.t(), phase)
    logmelmag2 = torch.log(mel_mag2 + cfg.safelog_eps)
    return torch.stack([logmelmag2, instantaneous_frequency(mel_phase)], dim=1)


def to_spectrogram(audio: torch.Tensor, cfg: FrontEndConfig = FrontEndConfig()
                   ) -> torch.Tensor:
    """``audio [B, T]`` -> ``[B, 2, n_freq, frames]`` in ``audio.dtype``."""
    spec = to_linear_spectrogram(audio, cfg)
    return linear_to_mel(spec, cfg) if cfg.use_mel_scale else spec


def epilogue(spec: torch.Tensor, masked_phase_threshold: Optional[float] = None,
             output_affine=None) -> torch.Tensor:
    """What the reference applies right after the transform: the masked-phase transform
    (IF := 0 where log-magnitude < threshold; extract_code.py:178-181) and then a per-channel
    affine normalisation ``((scale0, bias0), (scale1, bias1))`` (vqvae.py:254-255).  Both live
    in GANsynth_pytorch: unpinned, hence parameters."""
    out = spec.clone()
    if masked_phase_threshold is not None:
        out[:, 1] = torch.where(spec[:, 0] < masked_phase_threshold, torch.zeros_like(spec[:, 1]), spec[:, 1])
    if output_affine is not None:
        for c in range(2):
            out[:, c] = out[:, c] * output_affine[c][0] + output_affine[c][1]
    return out


def mel_to_linear_matrix(cfg: FrontEndConfig) -> np.ndarray:
    """float64 ``[n mel, n linear]`` approximate inverse of the filterbank (magenta
    ``specgrams_helper._mel_to_linear_matrix``): the transpose, each linear column divided
    by the column sum of ``M M^T`` (columns whose sum is ~0 keep that sum)."""
    m = linear_to_mel_matrix(cfg)
    sums = np.sum(np.matmul(m, m.T), axis=0)
    d = np.where(np.abs(sums) > 1.0e-8, 1.0 / np.where(sums == 0.0, 1.0, sums), sums)
    return np.matmul(m.T, np.diag(d))


def mel_to_linear(spec: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """magenta ``melspecgrams_to_specgrams``: ``[B, 2, n_freq, frames]`` mel log-mag^2 + mel IF
    -> linear log-magnitude + IF."""
    logmelmag2, mel_if = spec[:, 0], spec[:, 1]
    back = torch.from_numpy(mel_to_linear_matrix(cfg)).to(spec.dtype)     # [mel, lin]
    mag2 = torch.matmul(back.t(), torch.exp(logmelmag2))
    logmag = 0.5 * torch.log(mag2 + cfg.safelog_eps)
    mel_phase = torch.cumsum(mel_if * math.pi, dim=-1)
    phase = torch.matmul(back.t(), mel_phase)
    return torch.stack([logmag, instantaneous_frequency(phase)], dim=1)


def linear_to_stft(spec: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """magenta ``specgrams_to_stfts``: complex ``[B, n_fft/2 + 1, frames]``, the bin the
    forward transform dropped re-inserted as zero."""
    mag = torch.exp(spec[:, 0])
    phase = torch.cumsum(spec[:, 1] * math.pi, dim=-1)
    s = torch.polar(mag, phase)
    zero = torch.zeros_like(s[:, :1])
    return torch.cat([zero, s], dim=1) if cfg.drop_bin == "dc" else torch.cat([s, zero], dim=1)


def to_audio(spec: torch.Tensor, cfg: FrontEndConfig = FrontEndConfig(),
             input_affine=None) -> torch.Tensor:
    """``[B, 2, n_freq, frames]`` -> ``[B, hop * frames - pad_left]`` in ``spec.dtype``.

    Inverse of ``to_spectrogram`` (magenta ``melspecgrams_to_waves`` / ``specgrams_to_waves``):
    mel -> linear, IF -> phase by a running sum, inverse real FFT of every frame, synthesis
    window = analysis window, overlap-add divided by the summed squared windows of the frames
    that cover a sample, then the forward transform's padding removed (``n_fft - hop`` on the
    right).  ``input_affine`` ``((s0, b0), (s1, b1))`` is applied to the channels first (the
    shape of DataNormalizer.denormalize).  PARITY UNPINNED like the forward transform."""
    if input_affine is not None:
        spec = torch.stack([spec[:, c] * input_affine[c][0] + input_affine[c][1] for c in range(2)], dim=1)
    lin = mel_to_linear(spec, cfg) if cfg.use_mel_scale else spec
    s = linear_to_stft(lin, cfg)
    frames = s.shape[-1]
    pad_l = cfg.n_fft - cfg.hop_length if cfg.pad_left is None else cfg.pad_left
    total = cfg.hop_length * (frames - 1) + cfg.n_fft
    w = analysis_window(cfg, spec.dtype)
    cols = torch.fft.irfft(s, n=cfg.n_fft, dim=1) * w[None, :, None]            # [B, n_fft, frames]
    out = torch.zeros(spec.shape[0], total, dtype=spec.dtype)
    norm = torch.zeros(total, dtype=spec.dtype)
    for t in range(frames):
        out[:, t * cfg.hop_length:t * cfg.hop_length + cfg.n_fft] += cols[:, :, t]
        norm[t * cfg.hop_length:t * cfg.hop_length + cfg.n_fft] += w * w
    out = out / norm.clamp_min(1e-8)
    return out[:, pad_l:max(pad_l, total - (cfg.n_fft - cfg.hop_length))]


def stability_mask(audio: torch.Tensor, cfg: FrontEndConfig = FrontEndConfig(),
                   wrap_margin: float = 1e-2, mag_floor: float = 1e-4) -> torch.Tensor:
    """bool ``[B, n_freq, frames]``: True where the IF channel is numerically
    well-conditioned, evaluated in FP64.

    The IF is discontinuous where a wrapped phase step sits at +-pi, and the phase
    itself is ill-conditioned where |X| is tiny.  A (mel) bin/frame is flagged
    unstable when a contributing linear bin has |wrapped step| within
    ``wrap_margin`` of pi, or magnitude below ``mag_floor`` x the frame maximum in
    either frame of the step, or (mel) when the projected phase step is itself
    within ``wrap_margin`` of a wrap.  Parity tests compare IF only on stable
    positions and report how many were excluded.
    """
    a = audio.double()
    s = stft(a, cfg)
    mag = s.abs()
    ph = torch.angle(s)
    floor = mag.amax(dim=1, keepdim=True) * mag_floor
    small = mag < floor
    dd = _diff_time(ph)
    wrapped = torch.remainder(dd + math.pi, 2 * math.pi) - math.pi
    near = (math.pi - wrapped.abs()) < wrap_margin
    bad = torch.zeros_like(small)
    bad[..., 1:] = near | small[..., 1:] | small[..., :-1]
    bad[..., 0] = small[..., 0]
    if not cfg.use_mel_scale:
        return ~bad
    bank = torch.from_numpy(linear_to_mel_matrix(cfg))
    touched = torch.matmul((bank.t() > 0).double(), bad.double()) > 0
    lin = to_linear_spectrogram(a, cfg)
    mel_phase = torch.matmul(bank.t(), torch.cumsum(lin[:, 1] * math.pi, dim=-1))
    mdd = _diff_time(mel_phase)
    mwrapped = torch.remainder(mdd + math.pi, 2 * math.pi) - math.pi
    mnear = torch.zeros_like(touched)
    mnear[..., 1:] = (math.pi - mwrapped.abs()) < wrap_margin
    return ~(touched | mnear)
