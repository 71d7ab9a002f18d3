"""Two-level VQ-VAE-2 wrapper around the B200 quantiser: the *callers* of the hot path.

The conv encoder / decoder stay stock PyTorch (cuDNN), as in the reference; this file
only re-states their wiring so that code extraction can run ``encode`` alone
(the reference's ``extract_code.py:67`` runs the decoder and throws the
result away) and so that checkpoints of the reference load unchanged: module and
parameter names follow ``interactive_spectrogram_inpainting/vqvae/vqvae.py:130-218``
and ``encoder_decoder.py:18-227`` (``enc_b.blocks.N``, ``…conv.1``, ``quantize_conv_t``,
``quantize_t.embed`` …).

Scope notes: ``groups=1`` only; the fastai/xresnet encoders, the restart quantiser and
the GANSynth data normaliser (all external to the reference tree) are not covered.
"""
import json
import math
import pathlib
import threading
from typing import Mapping, Optional, Sequence, Type, Union

import torch
from torch import nn

from .. import _lib
from .bottleneck import EmaExchange, QuantizedBottleneck, UnquantizedBottleneck

# channel plan of the strided stages in quarters of `channel` (encoder_decoder.py:52-116)
_DOWN_PLAN = {16: (1, 2, 3, 4), 8: (2, 2, 4), 4: (2, 4), 2: (2,)}


# Inference-time epilogue fusion of the conv stacks.  A launch list of one extraction step
# (profiles/README.md, r02c) showed 49 % of the step in torch's separate bias-add and ReLU
# kernels around the cuDNN convolutions; under ``torch.no_grad()`` on CUDA the stacks below
# therefore call cuDNN's fused conv+bias(+residual)+ReLU (``torch.cudnn_convolution_relu`` /
# ``cudnn_convolution_add_relu``: library code, same math, no extra pass over the
# activations).  Training and CPU keep the stock modules.
fused_inference = True


def _can_fuse(x: torch.Tensor) -> bool:
    return (fused_inference and x.is_cuda and not torch.is_grad_enabled()
            and torch.backends.cudnn.enabled and not torch.is_autocast_enabled()
            and x.dtype in (torch.float32, torch.float16))


class TransposedFilters:
    """Filters for running a conv stack on the TRANSPOSED plane: a convolution commutes with
    swapping the two spatial axes if its kernel (and stride / padding pairs) are swapped too.
    The front end writes the spectrogram blocks frequency-fastest (``SpectrogramsHelper(
    space_to_depth="transposed")``: whole 128-byte lines per store); the encoder then reads them
    as ``[B, C, T, F]`` and uses ``w.transpose(2, 3)`` of every filter -- kept here, contiguous in
    channels_last, one copy per weight version."""

    def __init__(self):
        self._cache = {}

    def weight(self, w: torch.Tensor) -> torch.Tensor:
        key = (w._version, w.data_ptr(), w.device, w.dtype)
        hit = self._cache.get(id(w))
        if hit is not None and hit[0] == key:
            return hit[1]
        wt = w.detach().transpose(2, 3).contiguous(memory_format=torch.channels_last)
        self._cache[id(w)] = (key, wt)       # a racing thread stores an equal tensor
        return wt

    @staticmethod
    def pair(v):
        return tuple(reversed(v)) if isinstance(v, (tuple, list)) else v


def _w(conv: nn.Module, tf: Optional[TransposedFilters]) -> torch.Tensor:
    return conv.weight if tf is None else tf.weight(conv.weight)


def _hw(v, tf: Optional[TransposedFilters]):
    return v if tf is None else TransposedFilters.pair(v)


def _conv_relu(conv: nn.Conv2d, x: torch.Tensor, tf: Optional[TransposedFilters] = None) -> torch.Tensor:
    return torch.cudnn_convolution_relu(x, _w(conv, tf), conv.bias, _hw(conv.stride, tf),
                                        _hw(conv.padding, tf), _hw(conv.dilation, tf), conv.groups)


def _conv_add_relu(conv: nn.Conv2d, x: torch.Tensor, skip: torch.Tensor,
                   tf: Optional[TransposedFilters] = None) -> torch.Tensor:
    return torch.cudnn_convolution_add_relu(x, _w(conv, tf), skip, 1.0, conv.bias, _hw(conv.stride, tf),
                                            _hw(conv.padding, tf), _hw(conv.dilation, tf), conv.groups)


def _apply_transposed(m: nn.Module, x: torch.Tensor, tf: TransposedFilters) -> torch.Tensor:
    """``m(x)`` on the transposed plane, module by module (the stock ``forward`` would use the
    untransposed filters)."""
    F = torch.nn.functional
    if type(m) is nn.Conv2d:
        if m.padding_mode != 'zeros':
            raise NotImplementedError("transposed plane: zero padding only")
        return F.conv2d(x, tf.weight(m.weight), m.bias, tf.pair(m.stride), tf.pair(m.padding),
                        tf.pair(m.dilation), m.groups)
    if type(m) is nn.ConvTranspose2d:
        return F.conv_transpose2d(x, tf.weight(m.weight), m.bias, tf.pair(m.stride), tf.pair(m.padding),
                                  tf.pair(m.output_padding), m.groups, tf.pair(m.dilation))
    if isinstance(m, nn.ReLU):
        return torch.relu(x)
    if isinstance(m, ResBlock):
        x = torch.relu(x)
        return _apply_transposed(m.conv[3], torch.relu(_apply_transposed(m.conv[1], x, tf)), tf) + x
    raise NotImplementedError(f"transposed plane: no rule for {type(m).__name__}")


def _run_blocks(blocks: Sequence[nn.Module], x: torch.Tensor,
                tf: Optional[TransposedFilters] = None) -> torch.Tensor:
    """``blocks(x)``; with fusion, a conv absorbs the ReLU that follows it (explicit, or the
    one a ResBlock starts with), and a ResBlock absorbs the ReLU that follows IT (the next
    ResBlock's, or the stack's trailing one) -- so a ResBlock always sees a rectified input,
    which is also the value its skip connection reads (see ResBlock).  ``tf``: run on the
    transposed plane (``TransposedFilters``)."""
    if not _can_fuse(x):
        for m in blocks:
            x = m(x) if tf is None else _apply_transposed(m, x, tf)
        return x
    mods, i, rectified = list(blocks), 0, False
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        absorbs = isinstance(nxt, (nn.ReLU, ResBlock))
        if type(m) is nn.Conv2d and absorbs:
            x = _conv_relu(m, x, tf)
        elif isinstance(m, ResBlock) and absorbs:
            x = x if rectified else torch.relu(x)
            x = _conv_add_relu(m.conv[3], _conv_relu(m.conv[1], x, tf), x, tf)
        else:
            x, absorbs = (m(x) if tf is None else _apply_transposed(m, x, tf)), False
        rectified = absorbs or isinstance(m, nn.ReLU)
        i += 2 if (absorbs and isinstance(nxt, nn.ReLU)) else 1
    return x


def _as_rows(x: torch.Tensor):
    """``(row_stride)`` if ``x [B, C, H, W]`` is stored channels-last with uniformly strided
    rows ``(b, h, w)`` of ``C`` contiguous values, else None."""
    b, c, h, w = x.shape
    if x.stride(1) != 1 and c != 1:
        return None
    rs = x.stride(3) if w > 1 else (x.stride(2) if h > 1 else x.stride(0))
    if (w > 1 and x.stride(3) != rs) or (h > 1 and x.stride(2) != w * rs) or (b > 1 and x.stride(0) != h * w * rs):
        return None
    return rs if rs >= c and rs % 4 == 0 else None


class PointwiseProjection:
    """``conv(torch.cat(sources, 1))`` for a 1x1 ``Conv2d(C, 64)`` as ONE kernel of this repo
    (``isi_vq_project``, csrc/vq_project_tc.cu): the concatenation is never
    written, the bias is added in the epilogue, and the result comes back as the contiguous
    ``[B, H, W, 64]`` rows the nearest-code search reads (the reference permutes the NCHW conv
    output, vqvae.py:260,272).  ``folded_bias`` is the bias of a transposed convolution that
    produced ``sources[0]`` and was left out there: ``W[:, :c0] @ folded_bias`` joins the bias.

    Inference only (CUDA, ``no_grad``, channels-last sources with 64 | C, at least
    ``min_rows`` rows); ``usable`` says whether a call qualifies -- callers keep the stock modules
    otherwise."""

    min_rows = 128

    def __init__(self, conv: nn.Conv2d):
        self.conv = conv
        self._key = None
        self._prepared = None
        self._bias = None
        self._lock = threading.Lock()

    def usable(self, sources) -> bool:
        conv = self.conv
        if not (type(conv) is nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
                and conv.padding == (0, 0) and conv.groups == 1 and conv.out_channels == 64):
            return False
        if not 1 <= len(sources) <= 2 or not _can_fuse(sources[0]):
            return False
        if sum(x.shape[1] for x in sources) != conv.in_channels or conv.in_channels > 1024:
            return False
        shape = sources[0].shape
        for x in sources:
            if (x.dim() != 4 or x.dtype != torch.float32 or x.shape[1] % 64 or x.device != conv.weight.device
                    or x.shape[0] != shape[0] or x.shape[2:] != shape[2:] or _as_rows(x) is None
                    or x.data_ptr() % 16):
                return False
        return shape[0] * shape[2] * shape[3] >= self.min_rows

    def _prepare(self, c0: int, folded_bias):
        conv = self.conv
        w, b = conv.weight, conv.bias
        key = (w._version, w.data_ptr(), None if b is None else (b._version, b.data_ptr()),
               None if folded_bias is None else (folded_bias._version, folded_bias.data_ptr()), c0)
        if key == self._key:
            return self._prepared, self._bias
        with self._lock:
            return self._prepare_locked(key, c0, folded_bias)

    def _prepare_locked(self, key, c0: int, folded_bias):
        if key == self._key:
            return self._prepared, self._bias
        conv = self.conv
        w, b = conv.weight, conv.bias
        c_in = conv.in_channels
        w2 = w.detach().reshape(64, c_in).contiguous()
        nbytes = _lib.load().isi_vq_project_prepared_bytes(c_in)
        prepared = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        _lib.invoke("isi_vq_project_prepare", w2.data_ptr(), c_in, 64, prepared.data_ptr(), nbytes,
                    _lib.stream_ptr(w.device))
        bias = torch.zeros(64, device=w.device) if b is None else b.detach().clone()
        if folded_bias is not None:
            bias = bias + (w2[:, :c0].double() @ folded_bias.detach().double()).float()
        # fresh tensors per weight version: a concurrent caller keeps using the pair it was handed
        self._prepared, self._bias = prepared, bias.contiguous()
        self._key = key
        return self._prepared, self._bias

    def __call__(self, sources, folded_bias=None) -> torch.Tensor:
        x0 = sources[0]
        x1 = sources[1] if len(sources) > 1 else None
        prepared, bias = self._prepare(x0.shape[1], folded_bias)
        b, _, h, w = x0.shape
        out = torch.empty(b, h, w, 64, dtype=torch.float32, device=x0.device)
        _lib.invoke("isi_vq_project", x0.data_ptr(), x0.shape[1], _as_rows(x0),
                    None if x1 is None else x1.data_ptr(), 0 if x1 is None else x1.shape[1],
                    0 if x1 is None else _as_rows(x1), b * h * w, 64, prepared.data_ptr(),
                    bias.data_ptr(), out.data_ptr(), _lib.stream_ptr(x0.device))
        return out


class ResBlock(nn.Module):
    """relu -> 3x3 -> relu -> 1x1, added to the *rectified* input: the reference's leading
    in-place ReLU (encoder_decoder.py:22-35) rewrites the tensor the skip reads."""

    def __init__(self, in_channel: int, channel: int):
        super().__init__()
        self.conv = nn.Sequential(nn.ReLU(), nn.Conv2d(in_channel, channel, 3, padding=1),
                                  nn.ReLU(), nn.Conv2d(channel, in_channel, 1))

    def forward(self, x):
        x = torch.relu(x)
        return self.conv[3](torch.relu(self.conv[1](x))) + x


class Encoder(nn.Module):
    def __init__(self, in_channel: int, channel: int, n_res_block: int, n_res_channel: int,
                 resolution_factor: int, use_local_kernels: bool = False):
        super().__init__()
        if resolution_factor not in _DOWN_PLAN:
            raise ValueError(f"Unexpected resolution factor {resolution_factor}")
        k = 2 if use_local_kernels else 4
        widths = [q * channel // 4 for q in _DOWN_PLAN[resolution_factor]]
        blocks, prev = [], in_channel
        for w in widths:
            blocks += [nn.Conv2d(prev, w, k, stride=2, padding=1), nn.ReLU()]
            prev = w
        if resolution_factor == 2:           # single strided stage, then 3x3 widening
            blocks.append(nn.Conv2d(prev, channel, 3, padding=1))
        else:
            blocks.append(nn.Conv2d(channel, channel, 3, padding=1))
        blocks += [ResBlock(channel, n_res_channel) for _ in range(n_res_block)]
        blocks.append(nn.ReLU())
        self.blocks = nn.Sequential(*blocks)
        self._s2d_cache = None

    # (offset index d+1 in the 3x3 kernel, parity inside the 2x2 block, tap of the 4x4 kernel):
    # input row 2f'-1+k of the stride-2 conv is row parity p of block f'+d
    _S2D_TAPS = ((0, 1, 0), (1, 0, 1), (1, 1, 2), (2, 0, 3))

    def space_to_depth_weight(self) -> torch.Tensor:
        """The first convolution (4x4, stride 2, padding 1) as the equivalent 3x3 stride-1
        padding-1 kernel over the 2x2 space-to-depth input, ``[C_out, 4 C_in, 3, 3]`` with input
        channel = (f&1)*2C_in + (t&1)*C_in + c -- the layout the front-end kernel can write
        directly (``SpectrogramsHelper(space_to_depth=True)``).  Same products, regrouped."""
        conv = self.blocks[0]
        if not (isinstance(conv, nn.Conv2d) and conv.kernel_size == (4, 4) and conv.stride == (2, 2)
                and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1):
            raise ValueError("space_to_depth needs a 4x4 / stride 2 / padding 1 first convolution")
        w = conv.weight
        key = (w._version, w.device, w.data_ptr())
        if not torch.is_grad_enabled() and self._s2d_cache is not None and self._s2d_cache[0] == key:
            return self._s2d_cache[1]
        c_out, c_in = w.shape[:2]
        wp = w.new_zeros(c_out, 2, 2, c_in, 3, 3)
        for df, pf, kf in self._S2D_TAPS:
            for dt, pt, kt in self._S2D_TAPS:
                wp[:, pf, pt, :, df, dt] = w[:, :, kf, kt]
        wp = wp.reshape(c_out, 4 * c_in, 3, 3).contiguous(memory_format=torch.channels_last)
        if not torch.is_grad_enabled():
            self._s2d_cache = (key, wp)
        return wp

    def forward(self, x, space_to_depth: bool = False, transposed: Optional[TransposedFilters] = None):
        """``transposed``: ``x`` is on the transposed plane ``[B, C, T, F]`` (see TransposedFilters)."""
        if not space_to_depth:
            return _run_blocks(self.blocks, x, transposed)
        w, conv, rest = self.space_to_depth_weight(), self.blocks[0], list(self.blocks)[1:]
        if transposed is not None:
            w = transposed.weight(w)
        if _can_fuse(x) and rest and isinstance(rest[0], (nn.ReLU, ResBlock)):
            x = torch.cudnn_convolution_relu(x, w, conv.bias, (1, 1), (1, 1), (1, 1), 1)
            rest = rest[1:] if isinstance(rest[0], nn.ReLU) else rest
        else:
            x = torch.nn.functional.conv2d(x, w, conv.bias, 1, 1)
        return _run_blocks(rest, x, transposed)


class Decoder(nn.Module):
    def __init__(self, in_channel: int, out_channel: int, channel: int, n_res_block: int,
                 n_res_channel: int, resolution_factor: int, use_local_kernels: bool = False):
        super().__init__()
        if resolution_factor not in _DOWN_PLAN:
            raise ValueError(f"Unexpected resolution factor {resolution_factor}")
        k = 2 if use_local_kernels else 4
        blocks = [nn.Conv2d(in_channel, channel, 3, padding=1)]
        blocks += [ResBlock(channel, n_res_channel) for _ in range(n_res_block)]
        blocks.append(nn.ReLU())
        down = [q * channel // 4 for q in _DOWN_PLAN[resolution_factor]]
        # mirror of the encoder widths: channel -> ... -> out_channel
        ups = list(reversed(down[:-1])) + [out_channel] if resolution_factor != 2 else [out_channel]
        prev = channel
        for i, w in enumerate(ups):
            blocks.append(nn.ConvTranspose2d(prev, w, k, stride=2, padding=1))
            if i + 1 < len(ups):
                blocks.append(nn.ReLU())
            prev = w
        self.blocks = nn.Sequential(*blocks)

    def forward(self, x, without_last_bias: bool = False, transposed: Optional[TransposedFilters] = None):
        """``without_last_bias``: leave out the bias of the final transposed convolution (the
        caller folds it into the 1x1 projection that consumes the result, PointwiseProjection).
        ``transposed``: ``x`` is on the transposed plane (see TransposedFilters)."""
        last = self.blocks[-1]
        if not (without_last_bias and type(last) is nn.ConvTranspose2d and last.bias is not None):
            return _run_blocks(self.blocks, x, transposed)
        x = _run_blocks(list(self.blocks)[:-1], x, transposed)
        return torch.nn.functional.conv_transpose2d(
            x, _w(last, transposed), None, _hw(last.stride, transposed), _hw(last.padding, transposed),
            _hw(last.output_padding, transposed), last.groups, _hw(last.dilation, transposed))


class VQVAE(nn.Module):
    """``encode`` / ``decode`` / ``decode_code`` / ``forward`` with the reference's
    signatures (vqvae.py:245-302).  ``bottleneck_cls`` lets tests run the same wiring with
    the CPU oracle quantiser."""

    def __init__(self, in_channel: int = 3, num_hidden_channels: int = 128, n_res_block: int = 2,
                 num_residual_channels: int = 32, embed_dim: int = 64,
                 num_embeddings: Union[int, Sequence[int]] = 512, decay: float = 0.99,
                 groups: int = 1, use_local_kernels: bool = False,
                 resolution_factors: Mapping[str, int] = {'bottom': 4, 'top': 2},
                 embeddings_initial_variance: float = 1,
                 corruption_weights: Mapping[str, Optional[Sequence[float]]] = {'top': None,
                                                                             'bottom': None},
                 adapt_quantized_durations: bool = True,
                 output_activation_type: Optional[str] = None,
                 output_spectrogram_min_magnitude: Optional[float] = None,
                 decoder_output_activation: Optional[nn.Module] = None,
                 normalizer_statistics: Optional[Mapping[str, float]] = None,
                 disable_quantization: bool = False,
                 restarts_usage_threshold: float = 1.,
                 encoders=None, decoders=None,
                 bottleneck_cls: Optional[Type[nn.Module]] = None):
        """Every keyword of the reference constructor (vqvae.py:66-98) is named; the ones this
        wiring cannot honour raise instead of being dropped:

        * ``normalizer_statistics`` (vqvae.py:218-226): GANSynth's per-channel affine,
          ``{'s_a','s_b','p_a','p_b'}`` with ``normalised = x * a + b`` (``DataNormalizer`` lives
          in GANsynth_pytorch, which the reference does not vendor -- any other key set raises).
          ``encode`` normalises (vqvae.py:254-255), ``decode`` denormalises (vqvae.py:297-299).
        * ``output_spectrogram_min_magnitude`` (vqvae.py:238-241): the masked-phase output
          transform, IF := 0 where the log-magnitude is below the threshold (vqvae.py:300-301).
        * ``output_activation_type='threshold_gelu'`` (vqvae.py:228-236) only builds a module the
          reference never calls; accepted (with the reference's assert) and equally unused.
        * ``disable_quantization`` selects ``UnquantizedBottleneck`` (vqvae.py:159-160).
        * ``restarts_usage_threshold != 1`` (``QuantizedBottleneckWithRestarts``, built on the
          un-vendored ``discretization`` package), custom ``encoders`` / ``decoders`` (fastai /
          xresnet stacks), ``decoder_output_activation`` and ``groups != 1``: NotImplementedError.
        """
        super().__init__()
        if groups != 1:
            raise NotImplementedError("groups != 1")
        if decoder_output_activation is not None:
            raise NotImplementedError("decoder_output_activation (vqvae.py:99-100 raises too)")
        if encoders is not None or decoders is not None:
            raise NotImplementedError("custom encoders / decoders (fastai stacks) are out of scope")
        if restarts_usage_threshold != 1.:
            raise NotImplementedError("QuantizedBottleneckWithRestarts needs the un-vendored "
                                      "`discretization` package; restarts_usage_threshold must be 1")
        if output_activation_type not in (None, 'threshold_gelu'):
            raise AssertionError("Unexpected output activation type")          # vqvae.py:235-236
        if output_activation_type == 'threshold_gelu':
            assert output_spectrogram_min_magnitude is not None                  # vqvae.py:229
        if normalizer_statistics is not None:
            normalizer_statistics = dict(getattr(normalizer_statistics, '__dict__', normalizer_statistics))
            if set(normalizer_statistics) != {'s_a', 's_b', 'p_a', 'p_b'}:
                raise NotImplementedError(
                    "normalizer_statistics keys %s: only GANSynth's {'s_a','s_b','p_a','p_b'} affine "
                    "is implemented (GANsynth_pytorch.normalizer is not vendored by the reference)"
                    % sorted(normalizer_statistics))
        if bottleneck_cls is None:
            bottleneck_cls = UnquantizedBottleneck if disable_quantization else QuantizedBottleneck
        self._instantiation_parameters = dict(
            in_channel=in_channel, num_hidden_channels=num_hidden_channels,
            n_res_block=n_res_block, num_residual_channels=num_residual_channels,
            embed_dim=embed_dim, num_embeddings=num_embeddings, decay=decay, groups=groups,
            use_local_kernels=use_local_kernels, resolution_factors=dict(resolution_factors),
            embeddings_initial_variance=embeddings_initial_variance,
            corruption_weights=dict(corruption_weights),
            adapt_quantized_durations=adapt_quantized_durations,
            output_activation_type=output_activation_type,
            output_spectrogram_min_magnitude=output_spectrogram_min_magnitude,
            decoder_output_activation=None, normalizer_statistics=normalizer_statistics,
            disable_quantization=disable_quantization, restarts_usage_threshold=restarts_usage_threshold)
        self.output_activation_type = output_activation_type
        self.output_spectrogram_min_magnitude = output_spectrogram_min_magnitude
        self.normalizer_statistics = normalizer_statistics
        self.use_gansynth_normalization = normalizer_statistics is not None      # vqvae.py:215-216
        if normalizer_statistics is not None:
            st = normalizer_statistics
            self.register_buffer("_norm_scale", torch.tensor([st['s_a'], st['p_a']], dtype=torch.float32)
                                 .view(1, 2, 1, 1), persistent=False)
            self.register_buffer("_norm_bias", torch.tensor([st['s_b'], st['p_b']], dtype=torch.float32)
                                 .view(1, 2, 1, 1), persistent=False)
        self.in_channel, self.embed_dim = in_channel, embed_dim
        self.resolution_factors = dict(resolution_factors)
        self.adapt_quantized_durations = adapt_quantized_durations
        c, rb, rc = num_hidden_channels, n_res_block, num_residual_channels
        if isinstance(num_embeddings, int):
            self.n_embed_t = self.n_embed_b = num_embeddings
        else:
            self.n_embed_t, self.n_embed_b = num_embeddings

        self.enc_b = Encoder(in_channel, c, rb, rc, resolution_factors['bottom'], use_local_kernels)
        self.enc_t = Encoder(c, c, rb, rc, resolution_factors['top'], use_local_kernels)
        self.quantize_conv_t = nn.Conv2d(c, embed_dim, 1)
        self.quantize_t = bottleneck_cls(
            embed_dim, self.n_embed_t, decay=decay, corruption_weights=corruption_weights['top'],
            embeddings_initial_variance=embeddings_initial_variance)
        self.dec_t = Decoder(embed_dim, embed_dim, c, rb, rc, resolution_factors['top'],
                             use_local_kernels)
        self.quantize_conv_b = nn.Conv2d(embed_dim + c, embed_dim, 1)
        self.quantize_b = bottleneck_cls(
            embed_dim, self.n_embed_b, decay=decay,
            corruption_weights=corruption_weights['bottom'],
            embeddings_initial_variance=embeddings_initial_variance)
        k = 2 if use_local_kernels else 4
        self.upsample_top_to_bottom = nn.Sequential(*[
            nn.ConvTranspose2d(embed_dim, embed_dim, kernel_size=k, stride=2, padding=1)
            for _ in range(int(math.log2(resolution_factors['top'])))])
        self.dec = Decoder(2 * embed_dim, in_channel, c, rb, rc, resolution_factors['bottom'],
                           use_local_kernels)
        self._project_t = PointwiseProjection(self.quantize_conv_t)
        self._project_b = PointwiseProjection(self.quantize_conv_b)
        self._transposed_filters = TransposedFilters()
        # data-parallel training: ONE packed all-reduce of both quantisers' EMA statistics,
        # overlapped with the decoder; inert outside torch.distributed
        self.ema_exchange = (EmaExchange([self.quantize_t, self.quantize_b])
                             if isinstance(self.quantize_t, QuantizedBottleneck)
                             and type(self.quantize_t) is not UnquantizedBottleneck else None)

    # -- the two pre-quantiser 1x1 convolutions (vqvae.py:260 and :271-272), as ``[B, H, W, D]`` --
    def _prequant_top(self, enc_t: torch.Tensor) -> torch.Tensor:
        if self._project_t.usable([enc_t]):
            return self._project_t([enc_t])
        return self.quantize_conv_t(enc_t).permute(0, 2, 3, 1)

    def _prequant_bottom(self, quant_t: torch.Tensor, enc_b: torch.Tensor,
                         transposed: Optional[TransposedFilters] = None) -> torch.Tensor:
        last = self.dec_t.blocks[-1]
        fold = (not self.adapt_quantized_durations and type(last) is nn.ConvTranspose2d
                and last.bias is not None and _can_fuse(enc_b) and _as_rows(enc_b) is not None
                and enc_b.shape[1] % 64 == 0 and self.quantize_conv_b.out_channels == 64
                and enc_b.shape[0] * enc_b.shape[2] * enc_b.shape[3] >= PointwiseProjection.min_rows)
        dec_t = self.dec_t(quant_t, without_last_bias=fold, transposed=transposed)
        if self.adapt_quantized_durations:
            time_axis = 3 if transposed is None else 2
            n = min(dec_t.shape[time_axis], enc_b.shape[time_axis])
            dec_t, enc_b = dec_t.narrow(time_axis, 0, n), enc_b.narrow(time_axis, 0, n)
        if fold and self._project_b.usable([dec_t, enc_b]):
            return self._project_b([dec_t, enc_b], folded_bias=last.bias)
        if fold:
            dec_t = dec_t + last.bias.view(1, -1, 1, 1)
        if self._project_b.usable([dec_t, enc_b]):
            return self._project_b([dec_t, enc_b])
        return self.quantize_conv_b(torch.cat([dec_t, enc_b], 1)).permute(0, 2, 3, 1)

    # -- vqvae.py:251-278 --
    def encode(self, input: torch.Tensor, space_to_depth: bool = False, _exchange_open: bool = False):
        """``space_to_depth``: ``input`` is the 2x2-blocked spectrogram ``[B, 4C, F/2, T/2]`` that
        ``SpectrogramsHelper(space_to_depth=True)`` writes; same result, faster first conv."""
        ex = self.ema_exchange
        own_exchange = (self.training and ex is not None and not _exchange_open and input.is_cuda
                        and ex.begin(input.device))
        input = self._normalize(input, space_to_depth)
        enc_b = self.enc_b(input, space_to_depth=space_to_depth)
        enc_t = self.enc_t(enc_b)

        quant_t, diff_t, id_t, perplexity_t = self.quantize_t(self._prequant_top(enc_t))
        quant_t = quant_t.permute(0, 3, 1, 2)

        quant_b, diff_b, id_b, perplexity_b = self.quantize_b(self._prequant_bottom(quant_t, enc_b))
        quant_b = quant_b.permute(0, 3, 1, 2)
        if own_exchange:        # a bare encode() in training: exchange and update right away
            ex.launch()
            ex.finish()
        return (quant_t, quant_b, diff_t.unsqueeze(0) + diff_b.unsqueeze(0), id_t, id_b,
                perplexity_t, perplexity_b)

    def encode_codes(self, input: torch.Tensor, space_to_depth=False):
        """Top and bottom code maps only -- what ``extract_code.py`` stores.  Same codes as
        ``encode``; the bottom quantiser only searches (its lookup, commitment term and
        perplexity feed nothing here).

        ``space_to_depth="transposed"``: ``input`` is ``SpectrogramsHelper(space_to_depth=
        "transposed")``'s ``[B, 8, T/2, F/2]``; the whole encoder runs on the transposed plane
        (``TransposedFilters``) and the code maps are transposed back to ``[B, F', T']``."""
        if space_to_depth == "transposed":
            if self.training or not hasattr(self.quantize_b, "assign") or torch.is_grad_enabled():
                raise ValueError("space_to_depth='transposed' is an inference path: eval() and no_grad()")
            tf = self._transposed_filters
            input = self._normalize(input, True)
            enc_b = self.enc_b(input, space_to_depth=True, transposed=tf)
            enc_t = self.enc_t(enc_b, transposed=tf)
            quant_t, _, id_t, _ = self.quantize_t(self._prequant_top(enc_t))
            id_b = self.quantize_b.assign(self._prequant_bottom(quant_t.permute(0, 3, 1, 2), enc_b, tf))
            return id_t.transpose(1, 2), id_b.transpose(1, 2)
        if self.training or not hasattr(self.quantize_b, "assign"):
            out = self.encode(input, space_to_depth=space_to_depth)
            return out[3], out[4]
        input = self._normalize(input, space_to_depth)
        enc_b = self.enc_b(input, space_to_depth=space_to_depth)
        enc_t = self.enc_t(enc_b)
        quant_t, _, id_t, _ = self.quantize_t(self._prequant_top(enc_t))
        id_b = self.quantize_b.assign(self._prequant_bottom(quant_t.permute(0, 3, 1, 2), enc_b))
        return id_t, id_b

    # -- vqvae.py:254-255: DataNormalizer.normalize, unless the front end already applied it --
    normalizes_input = True

    def _normalize(self, input: torch.Tensor, space_to_depth: bool = False) -> torch.Tensor:
        if not (self.use_gansynth_normalization and self.normalizes_input):
            return input
        if space_to_depth:
            raise ValueError("space-to-depth input: let the front end apply the normaliser "
                             "(front_end_knobs()) and set model.normalizes_input = False")
        return input * self._norm_scale + self._norm_bias

    def front_end_knobs(self) -> dict:
        """The per-channel affine of ``encode`` as ``SpectrogramsHelper`` keywords: constructing the
        helper with ``output_affine=knobs['output_affine']`` fuses the normaliser into the front-end
        kernel's epilogue (then set ``model.normalizes_input = False``)."""
        if not self.use_gansynth_normalization:
            return {}
        st = self.normalizer_statistics
        return {"output_affine": ((st['s_a'], st['s_b']), (st['p_a'], st['p_b']))}

    def inverse_front_end_knobs(self) -> dict:
        """``post_process`` as keywords of the inverse front end (``to_audio``'s fused input
        affine): spectrogram = (decoder output - b) / a."""
        if not self.use_gansynth_normalization:
            return {}
        st = self.normalizer_statistics
        return {"input_affine": ((1.0 / st['s_a'], -st['s_b'] / st['s_a']),
                                 (1.0 / st['p_a'], -st['p_b'] / st['p_a']))}

    # -- vqvae.py:280-302 --
    def decode(self, quant_t: torch.Tensor, quant_b: torch.Tensor):
        dec = self.dec(torch.cat([self.upsample_top_to_bottom(quant_t), quant_b], 1))
        return self.post_process(dec)

    def post_process(self, dec: torch.Tensor) -> torch.Tensor:
        """vqvae.py:297-302: denormalise, then the masked-phase output transform."""
        if self.use_gansynth_normalization:
            dec = (dec - self._norm_bias) / self._norm_scale
        if self.output_spectrogram_min_magnitude is not None:
            keep = (dec[:, :1] >= self.output_spectrogram_min_magnitude).to(dec.dtype)
            dec = torch.cat([dec[:, :1], dec[:, 1:2] * keep, dec[:, 2:]], 1)
        return dec

    def decode_code(self, code_t: torch.Tensor, code_b: torch.Tensor):
        quant_t = self.quantize_t.embed_code(code_t).permute(0, 3, 1, 2)
        quant_b = self.quantize_b.embed_code(code_b).permute(0, 3, 1, 2)
        return self.decode(quant_t, quant_b)

    def forward(self, input):
        """vqvae.py:245-249.  In data-parallel training the packed EMA statistics of both
        quantisers are all-reduced while the decoder runs (``EmaExchange``)."""
        ex = self.ema_exchange
        exchanging = self.training and ex is not None and input.is_cuda and ex.begin(input.device)
        quant_t, quant_b, diff, id_t, id_b, perplexity_t, perplexity_b = self.encode(
            input, _exchange_open=exchanging)
        if exchanging:
            ex.launch()
        dec = self.decode(quant_t, quant_b)
        if exchanging:
            ex.finish()
        return dec, diff, perplexity_t, perplexity_b, id_t, id_b

    # -- vqvae.py:304-342 --
    @classmethod
    def from_parameters_and_weights(cls, parameters_json_path, model_weights_checkpoint_path,
                                    device: Union[str, torch.device] = 'cpu', **kwargs) -> 'VQVAE':
        with open(parameters_json_path, 'r') as f:
            parameters = json.load(f)
        model = cls(**parameters, **kwargs)
        state = torch.load(model_weights_checkpoint_path, map_location=device)
        if 'model' in state:
            state = state['model']
        model.load_state_dict(state, strict=False)
        return model

    def store_instantiation_parameters(self, path: pathlib.Path) -> None:
        with open(path, 'w') as f:
            json.dump(self._instantiation_parameters, f, indent=4)


class GraphedDecodeCode:
    """``model.decode_code`` for one fixed pair of code-map shapes, replayed from a CUDA graph.

    The interactive server decodes one edited pair of code maps per request
    (``flask_server.py:593-596``): at batch 1 the two lookups and the ~40 decoder kernels are
    launch-latency bound (0.46 ms eager, 0.20 ms replayed on a B200, profiles/README.md).  The
    graph captures the lookup kernels of this repo and the cuDNN decoder on a private stream;
    a call copies the new codes into the captured input buffers and replays.

    The model must be in eval mode with a codebook that no longer changes (the graph holds
    the address of the prepared codebook).  The returned tensor is the graph's output buffer:
    it is overwritten by the next call -- ``clone()`` it to keep it.  Not thread-safe; give
    each request thread its own instance.

    With ``to_audio`` (a spectrograms helper on the same device) the graph continues into the
    inverse front end (``isi_melif_inverse``), i.e. the server's whole codes -> audio request
    (``flask_server.py:593-596``: ``decode_code`` then ``spectrograms_helper.to_audio``), and a
    call returns ``(spectrogram, audio)``."""

    def __init__(self, model: VQVAE, code_t: torch.Tensor, code_b: torch.Tensor, warmup: int = 3,
                 to_audio=None):
        if model.training:
            raise RuntimeError("GraphedDecodeCode needs model.eval()")
        if not (code_t.is_cuda and code_b.is_cuda):
            raise RuntimeError("code maps must be CUDA tensors: there is no CPU fallback")
        self.model = model
        self._helper = to_audio

        def run():
            spec = model.decode_code(self._code_t, self._code_b)
            return spec if to_audio is None else (spec, to_audio.to_audio(spec))
        self._code_t = code_t.detach().long().clone()
        self._code_b = code_b.detach().long().clone()
        stream = torch.cuda.Stream(code_t.device)
        stream.wait_stream(torch.cuda.current_stream(code_t.device))
        with torch.no_grad(), torch.cuda.stream(stream):
            for _ in range(max(1, warmup)):        # cuDNN plan selection happens outside the capture
                run()
        torch.cuda.current_stream(code_t.device).wait_stream(stream)
        self._graph = torch.cuda.CUDAGraph()
        # thread_local: the server's other request threads keep using the CUDA runtime meanwhile
        with torch.no_grad(), torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            self._out = run()

    def __call__(self, code_t: torch.Tensor, code_b: torch.Tensor):
        if code_t.shape != self._code_t.shape or code_b.shape != self._code_b.shape:
            raise ValueError(f"captured for code maps {tuple(self._code_t.shape)} / "
                             f"{tuple(self._code_b.shape)}, got {tuple(code_t.shape)} / {tuple(code_b.shape)}")
        self._code_t.copy_(code_t, non_blocking=True)
        self._code_b.copy_(code_b, non_blocking=True)
        self._graph.replay()
        return self._out

