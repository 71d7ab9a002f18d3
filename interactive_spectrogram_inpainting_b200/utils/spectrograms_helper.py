"""``SpectrogramsHelper`` / ``MelSpectrogramsHelper`` backed by the fused sm_100a front end.

The reference builds these classes from the third-party ``GANsynth_pytorch`` package
(``interactive_spectrogram_inpainting/utils/misc.py:5-29``) whose source is not part of
the reference tree; constructor keywords and attribute names follow the reference's
call sites (``utils/misc.py:13-27``; ``fs_hz``/``hop_length``/``safelog_eps`` read at
``train_vqvae.py:400,421,711``; ``.to(device)`` at ``extract_code.py:173``).  The
arithmetic is the published GANSynth recipe; the choices the missing source would
decide are keyword-only arguments ("front end: unpinned").

``to_spectrogram`` is the hot path (isi_melif_forward).  ``to_audio`` is its mirror
(isi_melif_inverse); ``to_audio_differentiable`` is the same arithmetic in
plain torch for callers that need gradients (utils/losses/spectral.py:122-126).
"""
import math
from typing import Optional

import numpy as np
import torch
from torch import nn

from .. import _lib

_MEL_Q = 1127.0
SUPPORTED_N_FFT = (512, 1024, 2048)


_KERNEL_BAND_PITCH = 8     # kMaxMelWidth of csrc/melif_core.cuh


def mel_band_table(n_fft: int, fs_hz: float, lower_edge_hertz: float, upper_edge_hertz: float,
                   break_hz: float, width_factor: float):
    """Banded form of the GANSynth linear->mel filterbank for ``n_fft // 2`` linear and as
    many mel bins: for every mel bin the first linear row it touches, how many rows, and
    the FP64 triangle weights (zero-padded to the widest band).

    Mel scale m(f) = 1127 ln(1 + f/break); triangle edges equally spaced in mel between
    the two edge frequencies; a triangle narrower than ``width_factor`` linear bins is
    re-centred to that width in Hz (asinh form); linear bin frequencies are
    linspace(0, fs/2, n_bins) with the first row zeroed, as the public recipe has it.
    """
    n_bins = n_fft // 2
    nyquist = fs_hz / 2.0
    to_mel = lambda f: _MEL_Q * np.log1p(np.asarray(f, dtype=np.float64) / break_hz)
    to_hz = lambda m: break_hz * np.expm1(np.asarray(m, dtype=np.float64) / _MEL_Q)

    grid = np.linspace(to_mel(lower_edge_hertz), to_mel(upper_edge_hertz), n_bins + 2)
    left, centre, right = grid[:-2].copy(), grid[1:-1].copy(), grid[2:].copy()
    min_width_hz = width_factor * nyquist / n_bins
    narrow = (to_hz(right) - to_hz(left)) < min_width_hz
    half = _MEL_Q * np.arcsinh(0.5 * min_width_hz / (to_hz(centre) + break_hz))
    left = np.where(narrow, centre - half, left)
    right = np.where(narrow, centre + half, right)
    f_left, f_centre, f_right = to_hz(left), to_hz(centre), to_hz(right)

    bin_hz = np.linspace(0.0, nyquist, n_bins)
    starts = np.zeros(n_bins, dtype=np.int32)
    counts = np.zeros(n_bins, dtype=np.int32)
    bands = []
    for j in range(n_bins):
        lo = max(1, int(np.searchsorted(bin_hz, f_left[j], side="right")))
        hi = int(np.searchsorted(bin_hz, f_right[j], side="left"))      # exclusive
        f = bin_hz[lo:hi]
        w = np.minimum((f - f_left[j]) / (f_centre[j] - f_left[j]),
                       (f_right[j] - f) / (f_right[j] - f_centre[j]))
        keep = w > 0
        if keep.any():
            first = int(np.argmax(keep))
            last = len(keep) - int(np.argmax(keep[::-1]))
            starts[j], counts[j] = lo + first, last - first
            bands.append(np.maximum(w[first:last], 0.0))
        else:
            bands.append(np.zeros(0))
    width = max(1, int(counts.max()))
    weights = np.zeros((n_bins, width), dtype=np.float64)
    for j, b in enumerate(bands):
        weights[j, :len(b)] = b
    return starts, counts, weights


def dense_mel_matrix(starts, counts, weights) -> np.ndarray:
    """``[n linear, n mel]`` dense matrix equivalent to a band table."""
    n = len(starts)
    m = np.zeros((n, n), dtype=np.float64)
    for j in range(n):
        m[starts[j]:starts[j] + counts[j], j] = weights[j, :counts[j]]
    return m


def inverse_band_table(starts, counts, weights):
    """Banded form of GANSynth's mel->linear matrix (the transpose of the filterbank, every
    linear column divided by the column sum of ``M M^T``; columns whose sum is ~0 keep that
    sum): for every linear row the first mel row that feeds it, how many, FP64 weights."""
    m = dense_mel_matrix(starts, counts, weights)                  # [linear, mel]
    sums = (m @ m.T).sum(0)
    scale = np.where(np.abs(sums) > 1e-8, 1.0 / np.where(sums == 0.0, 1.0, sums), sums)
    back = m * scale[:, None]                                      # [linear, mel]: row l feeds from mel bins
    n = len(starts)
    inv_start = np.zeros(n, dtype=np.int32)
    inv_count = np.zeros(n, dtype=np.int32)
    bands = []
    for l in range(n):
        nz = np.flatnonzero(back[l])
        if len(nz):
            inv_start[l], inv_count[l] = nz[0], nz[-1] - nz[0] + 1
            bands.append(back[l, nz[0]:nz[-1] + 1])
        else:
            bands.append(np.zeros(0))
    width = max(1, int(inv_count.max()))
    inv_weight = np.zeros((n, width), dtype=np.float64)
    for l, b in enumerate(bands):
        inv_weight[l, :len(b)] = b
    return inv_start, inv_count, inv_weight


class SpectrogramsHelper(nn.Module):
    """Linear-frequency log-magnitude + instantaneous-frequency spectrograms."""

    use_mel_scale = False

    def __init__(self, fs_hz: int = 16000, n_fft: int = 2048, hop_length: int = 512,
                 window_length: int = 2048, safelog_eps: float = 1e-6, *,
                 pad_left: Optional[int] = None, n_frames: Optional[int] = None,
                 drop_bin: str = "dc", window_periodic: bool = True,
                 channels_last: bool = False, space_to_depth=False,
                 masked_phase_threshold: Optional[float] = None,
                 output_affine=None):
        super().__init__()
        if n_fft not in SUPPORTED_N_FFT:
            raise ValueError(f"n_fft must be one of {SUPPORTED_N_FFT}, got {n_fft}")
        if not 0 < window_length <= n_fft:
            raise ValueError("window_length must be in (0, n_fft]")
        if drop_bin not in ("dc", "nyquist"):
            raise ValueError("drop_bin must be 'dc' or 'nyquist'")
        self.fs_hz = fs_hz
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.window_length = window_length
        self.safelog_eps = safelog_eps
        self.pad_left = n_fft - hop_length if pad_left is None else pad_left
        self.fixed_n_frames = n_frames
        self.drop_bin = drop_bin
        # True: to_spectrogram returns the same [B,2,F,T'] tensor in torch.channels_last
        # storage, which is what the cuDNN conv encoder wants (no layout-conversion kernels)
        self.channels_last = channels_last
        # True: to_spectrogram returns the 2x2 space-to-depth form of the spectrogram,
        # ``[B, 8, F/2, T'/2]`` in channels_last storage with channel = (f&1)*4 + (t&1)*2 + c
        # (``from_space_to_depth`` undoes it).  The VQ-VAE's first convolution -- 4x4, stride 2
        # over 2 input channels, the slowest kernel of an extraction step -- is a 3x3 stride-1
        # convolution over these 8 channels (vqvae.Encoder, ``space_to_depth=True``).
        # "transposed": the same blocks on the TRANSPOSED plane, ``[B, 8, T'/2, F/2]`` in
        # channels_last storage (frequency runs fastest in memory): the layout the kernel's lanes
        # -- consecutive rows -- write as whole 128-byte lines, a quarter of the store
        # transactions of the time-fastest form.  ``VQVAE.encode_codes(x,
        # space_to_depth="transposed")`` runs the encoder on that plane with transposed filters.
        if space_to_depth not in (False, True, "transposed"):
            raise ValueError("space_to_depth must be False, True or 'transposed'")
        self.space_to_depth = space_to_depth
        # Fused epilogue (both are GANsynth_pytorch features the reference applies right after
        # the transform): the masked-phase transform -- IF := 0 where the log-magnitude is below
        # a threshold (extract_code.py:178-181, train_vqvae.py:586-589) -- and then a per-channel
        # affine ``((scale0, bias0), (scale1, bias1))``, the shape of DataNormalizer.normalize
        # (vqvae.py:254-255).  None = off.
        self.masked_phase_threshold = masked_phase_threshold
        self.output_affine = output_affine
        # int16 input only: the float value of one PCM step.  The dataset class the reference
        # uses (pytorch_nsynth, not in its tree) does this conversion on the CPU; its exact
        # constant is not pinned here, so it is an attribute.
        self.pcm_scale = 1.0 / 32768.0
        # to_audio: per-channel affine applied to the spectrogram first (None = off), and the
        # frames one CTA synthesises (None = chosen from the batch size)
        self.input_affine = None
        self.inverse_seg_frames = None

        w = torch.hann_window(window_length, periodic=window_periodic, dtype=torch.float64)
        if window_length < n_fft:
            lead = (n_fft - window_length) // 2
            w = torch.nn.functional.pad(w, (lead, n_fft - window_length - lead))
        self.register_buffer("window", w.float(), persistent=False)
        ang = -2.0 * math.pi * torch.arange(n_fft, dtype=torch.float64) / n_fft
        self.register_buffer("twiddle", torch.stack([ang.cos(), ang.sin()], dim=1).float(),
                             persistent=False)

    # ------------------------------------------------------------------
    @property
    def n_freq(self) -> int:
        return self.n_fft // 2

    def num_frames(self, n_samples: int) -> int:
        if self.fixed_n_frames is not None:
            return self.fixed_n_frames
        return max(1, math.ceil((n_samples + self.pad_left) / self.hop_length))

    def _params(self, n_frames: int) -> "_lib.MelifParams":
        p = _lib.MelifParams()
        p.n_fft, p.hop, p.pad_left, p.n_frames = self.n_fft, self.hop_length, self.pad_left, n_frames
        p.drop_dc = 1 if self.drop_bin == "dc" else 0
        p.use_mel, p.mel_width = 0, 0
        p.safelog_eps = self.safelog_eps
        p.window, p.twiddle = self.window.data_ptr(), self.twiddle.data_ptr()
        p.mel_start = p.mel_count = p.mel_weight = None
        p.channels_last = (_lib.SPEC_SPACE_TO_DEPTH_T if self.space_to_depth == "transposed" else
                           _lib.SPEC_SPACE_TO_DEPTH if self.space_to_depth else
                           _lib.SPEC_CHANNELS_LAST if self.channels_last else _lib.SPEC_PLANAR)
        p.mask_phase = 0 if self.masked_phase_threshold is None else 1
        p.mask_threshold = 0.0 if self.masked_phase_threshold is None else float(self.masked_phase_threshold)
        affine = self.output_affine or ((1.0, 0.0), (1.0, 0.0))
        for c in range(2):
            p.out_scale[c], p.out_bias[c] = float(affine[c][0]), float(affine[c][1])
        return p

    def to_spectrogram(self, audio: torch.Tensor) -> torch.Tensor:
        """``audio [B, T]`` (or ``[T]``) -> ``[B, 2, n_fft/2, frames]`` FP32 on the same GPU.

        ``int16`` audio is read as PCM, ``float(x) * self.pcm_scale`` (bit-identical to
        converting on the host first); every other dtype is taken as float samples."""
        _lib.require_cuda(audio, "audio")
        if audio.dim() == 1:
            audio = audio[None]
        if audio.dim() != 2:
            raise ValueError(f"audio must be [batch, samples], got {tuple(audio.shape)}")
        if self.window.device != audio.device:
            raise RuntimeError("helper and audio live on different devices; call .to(device)")
        a = audio.detach()
        if a.dtype != torch.int16 and a.dtype != torch.float32:
            a = a.float()
        if not a.is_contiguous():
            a = a.contiguous()
        n_notes, n_samples = a.shape
        frames = self.num_frames(n_samples)
        if self.hop_length * (frames - 1) + self.n_fft - n_samples - self.pad_left < 0:
            raise ValueError("n_frames too small for the audio length")
        if self.space_to_depth:
            if frames % 2 or self.n_freq % 2:
                raise ValueError("space_to_depth needs an even number of frames and bins")
            plane = ((frames // 2, self.n_freq // 2) if self.space_to_depth == "transposed"
                     else (self.n_freq // 2, frames // 2))
            out = torch.empty(n_notes, *plane, 8, dtype=torch.float32, device=a.device).permute(0, 3, 1, 2)
        else:
            out = torch.empty(n_notes, 2, self.n_freq, frames, dtype=torch.float32, device=a.device,
                              memory_format=(torch.channels_last if self.channels_last
                                             else torch.contiguous_format))
        if n_notes == 0:
            return out
        params = self._params(frames)
        if a.dtype == torch.int16:     # 16-bit PCM: converted by the kernel, half the upload
            params.audio_format, params.pcm_scale = _lib.AUDIO_PCM16, self.pcm_scale
        _lib.invoke("isi_melif_forward", a.data_ptr(), n_notes, n_samples, params,
                                                 out.data_ptr(), _lib.stream_ptr(a.device))
        return out

    forward = to_spectrogram

    @staticmethod
    def from_space_to_depth(blocks: torch.Tensor, transposed: bool = False) -> torch.Tensor:
        """``[B, 8, F/2, T/2]`` (channel = (f&1)*4 + (t&1)*2 + c) -> ``[B, 2, F, T]``;
        ``transposed``: from the ``[B, 8, T/2, F/2]`` form."""
        if transposed:
            blocks = blocks.transpose(2, 3)
        b, c8, f2, t2 = blocks.shape
        x = blocks.reshape(b, 2, 2, c8 // 4, f2, t2)            # [B, pf, pt, c, F/2, T/2]
        return x.permute(0, 3, 4, 1, 5, 2).reshape(b, c8 // 4, 2 * f2, 2 * t2)

    @staticmethod
    def to_space_to_depth(spec: torch.Tensor) -> torch.Tensor:
        """``[B, C, F, T]`` -> ``[B, 4C, F/2, T/2]`` with channel = (f&1)*2C + (t&1)*C + c."""
        b, c, f, t = spec.shape
        x = spec.reshape(b, c, f // 2, 2, t // 2, 2)            # [B, c, F/2, pf, T/2, pt]
        return x.permute(0, 3, 5, 1, 2, 4).reshape(b, 4 * c, f // 2, t // 2)

    # ------------------------------------------------------------------
    # Inverse (SURVEY.md 8f N4 -- callers of the hot path: flask_server.py:596,1016,1110,
    # sample.py:599, train_vqvae.py:392-394, utils/losses/spectral.py:122-126).
    # ------------------------------------------------------------------
    def _ola_scale(self, frames: int, device) -> torch.Tensor:
        """``1 / (n_fft * sum of squared windows of the frames covering a sample)`` for every
        padded position of a ``frames``-frame inverse STFT (FP64 on the host, cached)."""
        key = (frames, str(device))
        cache = self.__dict__.setdefault("_ola_cache", {})
        if key not in cache:
            w2 = self.window.double().cpu() ** 2
            total = self.hop_length * (frames - 1) + self.n_fft
            norm = torch.zeros(total, dtype=torch.float64)
            for t in range(frames):
                norm[t * self.hop_length:t * self.hop_length + self.n_fft] += w2
            cache[key] = (1.0 / (self.n_fft * norm.clamp_min(1e-8))).float().to(device)
        return cache[key]

    def _inverse_params(self, frames: int, device) -> "_lib.ImelifParams":
        p = _lib.ImelifParams()
        p.n_fft, p.hop, p.pad_left, p.n_frames = self.n_fft, self.hop_length, self.pad_left, frames
        p.drop_dc = 1 if self.drop_bin == "dc" else 0
        p.use_mel, p.band_width = 0, 0
        p.safelog_eps = self.safelog_eps
        p.window, p.twiddle = self.window.data_ptr(), self.twiddle.data_ptr()
        p.band_start = p.band_count = p.band_weight = None
        p.ola_scale = self._ola_scale(frames, device).data_ptr()
        affine = self.input_affine or ((1.0, 0.0), (1.0, 0.0))
        for c in range(2):
            p.in_scale[c], p.in_bias[c] = float(affine[c][0]), float(affine[c][1])
        p.seg_frames = int(self.inverse_seg_frames or 0)
        return p

    def to_audio(self, spec: torch.Tensor) -> torch.Tensor:
        """``[B, 2, n_fft/2, frames]`` (the layout ``to_spectrogram`` returns) -> ``[B, hop *
        frames - pad_left]`` FP32 on the same GPU: mel -> linear, IF -> phase by a running sum,
        inverse STFT with the analysis window, the forward padding removed.  ``self.input_affine``
        ``((s0, b0), (s1, b1))`` is applied to the channels first (DataNormalizer.denormalize's
        shape).  No gradient: see ``to_audio_differentiable``."""
        _lib.require_cuda(spec, "spec")
        if spec.dim() == 3:
            spec = spec[None]
        if spec.dim() != 4 or spec.shape[1] != 2 or spec.shape[2] != self.n_freq:
            raise ValueError(f"spec must be [batch, 2, {self.n_freq}, frames], got {tuple(spec.shape)}")
        if self.window.device != spec.device:
            raise RuntimeError("helper and spectrogram live on different devices; call .to(device)")
        if spec.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("to_audio does not record gradients; use to_audio_differentiable")
        x = spec.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        n_notes, frames = x.shape[0], x.shape[3]
        n_samples = self.hop_length * frames - self.pad_left
        if frames == 0 or n_samples <= 0:
            raise ValueError("spectrogram too short for the helper's padding")
        out = torch.empty(n_notes, n_samples, dtype=torch.float32, device=x.device)
        if n_notes == 0:
            return out
        params = self._inverse_params(frames, x.device)
        _lib.invoke("isi_melif_inverse", x.data_ptr(), n_notes, params, out.data_ptr(), n_samples,
                    _lib.stream_ptr(x.device))
        return out

    def _linear_to_audio(self, spec: torch.Tensor) -> torch.Tensor:
        """``[B, 2, F, T']`` linear log-magnitude + IF -> ``[B, samples]`` (GANSynth
        ``specgrams_to_stfts`` + inverse STFT, padding removed), plain torch."""
        logmag, ifreq = spec[:, 0].float(), spec[:, 1].float()
        mag = torch.exp(logmag)
        phase = torch.cumsum(ifreq * math.pi, dim=-1)
        stft = torch.polar(mag, phase)
        zero = torch.zeros_like(stft[:, :1])
        stft = torch.cat([zero, stft], 1) if self.drop_bin == "dc" else torch.cat([stft, zero], 1)
        frames = stft.shape[-1]
        total = self.hop_length * (frames - 1) + self.n_fft
        window = self.window.to(stft.device)
        # overlap-add by hand: torch.istft insists on center=True framing
        cols = torch.fft.irfft(stft, n=self.n_fft, dim=1) * window[None, :, None]
        audio = torch.nn.functional.fold(cols, (1, total), (1, self.n_fft), stride=(1, self.hop_length))
        norm = torch.nn.functional.fold((window ** 2)[None, :, None].expand(1, -1, frames).contiguous(),
                                        (1, total), (1, self.n_fft), stride=(1, self.hop_length))
        audio = (audio / norm.clamp_min(1e-8))[:, 0, 0]
        pad_right = self.n_fft - self.hop_length
        return audio[:, self.pad_left:total - pad_right]

    def _denormalised(self, spec: torch.Tensor) -> torch.Tensor:
        if self.input_affine is None:
            return spec
        return torch.stack([spec[:, c] * self.input_affine[c][0] + self.input_affine[c][1]
                            for c in range(2)], 1)

    def to_audio_differentiable(self, spec: torch.Tensor) -> torch.Tensor:
        """``to_audio`` as differentiable torch ops (any device), for the spectral training
        losses of the reference (utils/losses/spectral.py:122-126)."""
        return self._linear_to_audio(self._denormalised(spec))

    def from_wavfile(self, path, duration_n: Optional[int] = None) -> torch.Tensor:
        """Load a wav file (mono mix, resampled to ``fs_hz``, cropped / zero-padded to
        ``duration_n`` samples) and return its ``[1, 2, F, T']`` spectrogram on the helper's
        device (call shape of flask_server.py:648, sample.py:526)."""
        import torchaudio
        wav, fs = torchaudio.load(str(path))
        wav = wav.mean(0, keepdim=True)
        if fs != self.fs_hz:
            wav = torchaudio.functional.resample(wav, fs, self.fs_hz)
        if duration_n is not None:
            wav = wav[:, :duration_n]
            if wav.shape[1] < duration_n:
                wav = torch.nn.functional.pad(wav, (0, duration_n - wav.shape[1]))
        return self.to_spectrogram(wav.to(self.window.device))


class MelSpectrogramsHelper(SpectrogramsHelper):
    """Mel-scaled variant: log of the mel-projected squared magnitude and the IF of the
    mel-projected unwrapped phase (GANSynth ``specgrams_to_melspecgrams``)."""

    use_mel_scale = True

    def __init__(self, fs_hz: int = 16000, n_fft: int = 2048, hop_length: int = 512,
                 window_length: int = 2048, safelog_eps: float = 1e-6,
                 lower_edge_hertz: float = 0.0, upper_edge_hertz: float = 8000.0,
                 mel_break_frequency_hertz: float = 700.0,
                 mel_bin_width_threshold_factor: float = 1.5, **knobs):
        super().__init__(fs_hz, n_fft, hop_length, window_length, safelog_eps, **knobs)
        self.lower_edge_hertz = lower_edge_hertz
        self.upper_edge_hertz = upper_edge_hertz
        self.mel_break_frequency_hertz = mel_break_frequency_hertz
        self.mel_bin_width_threshold_factor = mel_bin_width_threshold_factor
        starts, counts, weights = mel_band_table(
            n_fft, fs_hz, lower_edge_hertz, upper_edge_hertz, mel_break_frequency_hertz,
            mel_bin_width_threshold_factor)
        inv_start, inv_count, inv_weight = inverse_band_table(starts, counts, weights)
        if inv_weight.shape[1] < _KERNEL_BAND_PITCH:
            inv_weight = np.pad(inv_weight, ((0, 0), (0, _KERNEL_BAND_PITCH - inv_weight.shape[1])))
        self.register_buffer("inv_start", torch.from_numpy(inv_start), persistent=False)
        self.register_buffer("inv_count", torch.from_numpy(inv_count), persistent=False)
        self.register_buffer("inv_weight", torch.from_numpy(inv_weight).float().contiguous(), persistent=False)
        self.register_buffer("mel_start", torch.from_numpy(starts), persistent=False)
        self.register_buffer("mel_count", torch.from_numpy(counts), persistent=False)
        if weights.shape[1] < _KERNEL_BAND_PITCH:      # 32-byte rows: two 16-byte loads per band
            weights = np.pad(weights, ((0, 0), (0, _KERNEL_BAND_PITCH - weights.shape[1])))
        self.register_buffer("mel_weight", torch.from_numpy(weights).float().contiguous(),
                             persistent=False)

    def to_audio_differentiable(self, spec: torch.Tensor) -> torch.Tensor:
        """GANSynth ``melspecgrams_to_specgrams`` (pseudo-inverse filterbank: transpose
        normalised by the column sums of M M^T) followed by the linear inverse, plain torch."""
        spec = self._denormalised(spec)
        logmelmag2, mel_if = spec[:, 0].float(), spec[:, 1].float()
        m = torch.from_numpy(dense_mel_matrix(self.mel_start.cpu().numpy(), self.mel_count.cpu().numpy(),
                                              self.mel_weight.double().cpu().numpy())).to(spec.device)
        gram_rows = (m @ m.t()).sum(0)
        scale = torch.where(gram_rows.abs() > 1e-8, 1.0 / gram_rows, gram_rows)
        mel_to_lin = (m.t() * scale[None, :]).float()                  # [mel, linear]
        mag2 = torch.einsum("bmt,ml->blt", torch.exp(logmelmag2), mel_to_lin)
        logmag = 0.5 * torch.log(mag2.clamp_min(0) + self.safelog_eps)
        mel_phase = torch.cumsum(mel_if * math.pi, dim=-1)
        phase = torch.einsum("bmt,ml->blt", mel_phase, mel_to_lin)
        ifreq = torch.cat([phase[..., :1], phase[..., 1:] - phase[..., :-1]], -1) / math.pi
        return self._linear_to_audio(torch.stack([logmag, ifreq], 1))

    def _inverse_params(self, frames: int, device) -> "_lib.ImelifParams":
        p = super()._inverse_params(frames, device)
        p.use_mel, p.band_width = 1, self.inv_weight.shape[1]
        p.band_start, p.band_count = self.inv_start.data_ptr(), self.inv_count.data_ptr()
        p.band_weight = self.inv_weight.data_ptr()
        return p

    def _params(self, n_frames: int) -> "_lib.MelifParams":
        p = super()._params(n_frames)
        p.use_mel, p.mel_width = 1, self.mel_weight.shape[1]
        p.mel_start, p.mel_count = self.mel_start.data_ptr(), self.mel_count.data_ptr()
        p.mel_weight = self.mel_weight.data_ptr()
        return p
