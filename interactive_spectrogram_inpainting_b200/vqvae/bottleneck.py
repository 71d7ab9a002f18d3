"""Drop-in ``QuantizedBottleneck`` backed by the sm_100a kernels.

Mirrors the reference class
``interactive_spectrogram_inpainting/vqvae/bottleneck.py:30-104`` -- constructor
signature (:33-36), buffer names/shapes ``embed [D,K]``, ``cluster_size [K]``,
``embed_avg [D,K]`` (:49-51, so checkpoints load with the same keys), the 4-tuple
returned by ``forward`` (:53-54,101) and ``embed_code`` (:103-104) -- while the
arithmetic runs in ``libisi_b200.so``:

    isi_vq_assign        distance + argmin            (bottleneck.py:55-61)
    isi_vq_gather_stats  lookup, (q-x)^2, EMA sums    (bottleneck.py:75-83,94-95)
    isi_vq_finish        diff and perplexity scalars  (bottleneck.py:94,97-100)
    isi_vq_ema_update    EMA blend + normalise        (bottleneck.py:80-92)

There is no CPU path: CPU tensors raise.
"""
import math
from typing import List, Optional, Sequence, Tuple

import threading

import torch
from torch import nn

from .. import _lib


class _CodebookCache:
    """Device scratch derived from ``embed`` (isi_vq_prepare_codebook), refreshed when the
    buffer is written (EMA step, ``load_state_dict``, ``.to()``)."""

    def __init__(self):
        self.buffer = None
        self.key = None
        self._lock = threading.Lock()     # request threads of the Flask server share the module

    def get(self, embed: torch.Tensor) -> torch.Tensor:
        key = (embed.data_ptr(), embed._version, embed.device, tuple(embed.shape))
        if self.buffer is not None and self.key == key:
            return self.buffer
        with self._lock:
            if self.buffer is not None and self.key == key:
                return self.buffer
            lib = _lib.load()
            dim, n_embed = embed.shape
            nbytes = lib.isi_vq_prepared_bytes(dim, n_embed)
            if (self.buffer is None or self.buffer.numel() < nbytes
                    or self.buffer.device != embed.device):
                self.buffer = torch.empty(nbytes, dtype=torch.uint8, device=embed.device)
            src = embed if embed.is_contiguous() else embed.contiguous()
            _lib.invoke("isi_vq_prepare_codebook",
                        src.data_ptr(), dim, n_embed, self.buffer.data_ptr(), self.buffer.numel(),
                        _lib.stream_ptr(embed.device))
            self.key = key
            return self.buffer

    def invalidate(self):
        self.key = None


def _as_rows(t: torch.Tensor) -> Tuple[torch.Tensor, "_lib.RowsLayout"]:
    layout = _lib.rows_layout(t)
    if layout is None:
        t = t.contiguous()
        layout = _lib.rows_layout(t)
    return t, layout


def _empty_like_strided(t: torch.Tensor) -> torch.Tensor:
    """Output with the strides the reference result has: ``input + (...)`` keeps the
    layout of ``input`` (a permuted NCHW view at vqvae.py:260,272 -> permuting back gives
    a contiguous NCHW tensor)."""
    if t.numel() == 0 or t.is_contiguous():
        return torch.empty_like(t, memory_format=torch.contiguous_format)
    return torch.empty_like(t, memory_format=torch.preserve_format)


class _QuantizeFunction(torch.autograd.Function):
    """Straight-through estimator of bottleneck.py:94-95: d quantize / d input = I and
    d diff / d input = 2 (x - q) / (N D); the codebook gets no gradient (EMA)."""

    @staticmethod
    def forward(ctx, module: "QuantizedBottleneck", x: torch.Tensor):
        quantize, diff, ind, perplexity = module._forward_impl(x)
        ctx.save_for_backward(x, quantize)
        ctx.mark_non_differentiable(ind, perplexity)
        return quantize, diff, ind, perplexity

    @staticmethod
    def backward(ctx, g_quantize, g_diff, _g_ind, _g_perp):
        x, quantize = ctx.saved_tensors
        grad = g_quantize
        if g_diff is not None:
            scale = g_diff * (2.0 / x.numel())
            extra = (x - quantize) * scale
            grad = extra if grad is None else grad + extra
        return None, grad


class EmaExchange:
    """The training path's one exchange step: the EMA statistics of SEVERAL
    quantisers (the VQ-VAE's top and bottom) live in ONE packed buffer

        [counts_t (K_t) | embed_sum_t (K_t*D) | counts_b (K_b) | embed_sum_b (K_b*D)]      266 KB

    which is summed over the ranks by a single ``all_reduce(async_op=True)`` issued right after
    the last quantiser ran -- NCCL over NVLink on its own stream -- and waited for only when the
    codebook updates are applied, i.e. after whatever the caller enqueued in between (the
    decoder's forward in ``VQVAE.forward``).  Every rank then runs the identical
    ``isi_vq_ema_update`` on identical sums, so the codebooks stay bit-identical across ranks.

    Protocol (driven by the owner): ``begin()`` -> each member quantiser's training forward
    writes its statistics into ``slot(member)`` and defers its update -> ``launch()`` ->
    ... -> ``finish()``.  Inactive (members update immediately, as a lone quantiser does) unless
    torch.distributed is initialised with more than one rank and every member syncs."""

    def __init__(self, members: Sequence["QuantizedBottleneck"], process_group=None):
        self.members = list(members)
        self.process_group = process_group
        self.offsets, off = [], 0
        for m in self.members:
            self.offsets.append(off)
            off += m.n_embed * (1 + m.dim)
        self.numel = off
        self.pack: Optional[torch.Tensor] = None
        self.open = False
        self._pending: List["QuantizedBottleneck"] = []
        self._work = None
        self.last_allreduce_events = None      # (start, end) CUDA events of the last collective, if timed
        self.time_collective = False

    def enabled(self) -> bool:
        import torch.distributed as dist
        return (dist.is_available() and dist.is_initialized()
                and dist.get_world_size(self.process_group) > 1
                and all(m.sync_ema_stats and m.training for m in self.members))

    def begin(self, device: torch.device) -> bool:
        if not self.enabled():
            return False
        if self.pack is None or self.pack.device != device:
            self.pack = torch.zeros(self.numel, dtype=torch.float32, device=device)
        else:
            self.pack.zero_()
        self.open, self._pending, self._work = True, [], None
        for m in self.members:
            m._exchange = self
        return True

    def slot(self, member: "QuantizedBottleneck") -> torch.Tensor:
        i = next(k for k, m in enumerate(self.members) if m is member)
        return self.pack[self.offsets[i]: self.offsets[i] + member.n_embed * (1 + member.dim)]

    def defer(self, member: "QuantizedBottleneck") -> None:
        self._pending.append(member)

    def launch(self) -> None:
        import torch.distributed as dist
        if not self.open:
            return
        if self.time_collective:
            start = torch.cuda.Event(enable_timing=True)
            start.record()
        self._work = dist.all_reduce(self.pack, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
        if self.time_collective:
            self._start = start

    def finish(self) -> None:
        if not self.open:
            return
        if self._work is not None:
            self._work.wait()                  # the current stream waits for the collective
            if self.time_collective:
                end = torch.cuda.Event(enable_timing=True)
                end.record()
                self.last_allreduce_events = (self._start, end)
        for m in self._pending:
            m._apply_ema(self.slot(m))
        self.open, self._pending, self._work = False, [], None
        for m in self.members:
            m._exchange = None


class QuantizedBottleneck(nn.Module):
    cluster_size: torch.Tensor

    def __init__(self, dim: int, n_embed: int, decay: float = 0.99, eps: float = 1e-5,
                 embeddings_initial_variance: float = 1,
                 corruption_weights: Optional[List[float]] = None):
        super().__init__()
        self.dim = dim
        self.n_embed = n_embed
        self.decay = decay
        self.eps = eps
        self.corruption_weights = corruption_weights
        self.embeddings_initial_variance = embeddings_initial_variance

        embed = torch.randn(dim, n_embed) * math.sqrt(self.embeddings_initial_variance)
        self.register_buffer('embed', embed)
        self.register_buffer('cluster_size', torch.zeros(n_embed))
        self.register_buffer('embed_avg', embed.clone())

        # knobs that do not exist in the reference
        self.assign_algo = "auto"              # "auto" | "simt" | "tcgen05"
        self.sync_ema_stats = True             # all-reduce EMA sums when torch.distributed is up
        self.stats_process_group = None
        self.embed_code_channels_first = True  # 3-D ids -> NCHW storage (vqvae.py:289-292)
        self._cache = _CodebookCache()
        self._exchange: Optional[EmaExchange] = None   # set by an open EmaExchange for one step

    # -- nn.Module plumbing: any re-materialisation of the buffers drops the cache --
    def _apply(self, fn, *args, **kwargs):
        self._cache.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._cache.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------
    def assign(self, input: torch.Tensor) -> torch.Tensor:
        """Nearest-code indices only (bottleneck.py:55-61), shape ``input.shape[:-1]``."""
        x, layout, n_rows = self._check_input(input)
        ind = torch.empty(n_rows, dtype=torch.int64, device=x.device)
        if n_rows:
            _lib.invoke("isi_vq_assign", 
                x.data_ptr(), layout, n_rows, self.dim, self.n_embed,
                self._cache.get(self.embed).data_ptr(), ind.data_ptr(), None,
                _lib.algo_id(self.assign_algo), _lib.stream_ptr(x.device))
        return ind.view(*input.shape[:-1])

    def forward(self, input: torch.Tensor):
        if torch.is_grad_enabled() and input.requires_grad:
            return _QuantizeFunction.apply(self, input)
        return self._forward_impl(input)

    def embed_code(self, embed_id: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(embed_id, "embed_id")
        if embed_id.dtype != torch.int64:
            embed_id = embed_id.long()
        ids = embed_id if embed_id.is_contiguous() else embed_id.contiguous()
        n_rows = ids.numel()
        if ids.dim() == 3 and self.embed_code_channels_first:
            b, h, w = ids.shape
            storage = torch.empty(b, self.dim, h, w, dtype=torch.float32, device=ids.device)
            out = storage.permute(0, 2, 3, 1)
            layout = _lib.RowsLayout(h * w, self.dim * h * w, 1, h * w)
        else:
            out = torch.empty(*ids.shape, self.dim, dtype=torch.float32, device=ids.device)
            layout = _lib.RowsLayout(max(n_rows, 1), 0, self.dim, 1)
        if n_rows:
            flag = torch.zeros(1, dtype=torch.int32, device=ids.device)
            _lib.invoke("isi_embed_code", 
                ids.data_ptr(), n_rows, self.dim, self.n_embed,
                self._cache.get(self.embed).data_ptr(), out.data_ptr(), layout,
                flag.data_ptr(), _lib.stream_ptr(ids.device))
            torch._assert_async(flag[0] == 0, "embed_code: index out of range")
        return out

    # ------------------------------------------------------------------
    def _check_input(self, input: torch.Tensor):
        _lib.require_cuda(input, "input")
        if input.shape[-1] != self.dim:
            raise ValueError(f"last dimension must be {self.dim}, got {tuple(input.shape)}")
        if self.embed.device != input.device:
            raise RuntimeError("codebook and input live on different devices")
        x = input.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x, layout = _as_rows(x)
        return x, layout, x.numel() // self.dim

    def _forward_impl(self, input: torch.Tensor):
        lib = _lib.load()
        x, layout, n_rows = self._check_input(input)
        dev, stream = x.device, _lib.stream_ptr(x.device)
        if n_rows == 0:
            raise ValueError("empty input")
        prepared = self._cache.get(self.embed)

        ind = torch.empty(n_rows, dtype=torch.int64, device=dev)
        _lib.invoke("isi_vq_assign", x.data_ptr(), layout, n_rows, self.dim, self.n_embed,
                                     prepared.data_ptr(), ind.data_ptr(), None,
                                     _lib.algo_id(self.assign_algo), stream)

        if self.training and self.corruption_weights is not None:
            # bottleneck.py:63-73 (the reference draws on the CPU; same distribution)
            weights = torch.tensor(self.corruption_weights, dtype=torch.float32, device=dev)
            shift = torch.multinomial(weights, n_rows, replacement=True) - 1
            ind = (ind + shift) % self.n_embed

        quantize = _empty_like_strided(x)
        q_layout = _lib.rows_layout(quantize)
        n_stats = self.n_embed * (1 + self.dim) if self.training else self.n_embed
        deferred = self.training and self._exchange is not None and self._exchange.open
        # an open exchange lends (zeroed) room in its packed buffer; the update is its job
        stats = self._exchange.slot(self) if deferred else torch.zeros(n_stats, dtype=torch.float32, device=dev)
        ws_bytes = lib.isi_vq_gather_workspace_bytes(n_rows, self.dim)
        workspace = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=dev)
        _lib.invoke("isi_vq_gather_stats", 
            x.data_ptr(), layout, ind.data_ptr(), n_rows, self.dim, self.n_embed,
            prepared.data_ptr(), quantize.data_ptr(), q_layout, stats.data_ptr(),
            0 if self.training else 1, workspace.data_ptr(), workspace.numel() * 8, None, stream)

        scalars = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.invoke("isi_vq_finish", workspace.data_ptr(), n_rows, self.dim, self.n_embed,
                                     stats.data_ptr(), scalars.data_ptr(),
                                     scalars.data_ptr() + 4, stream)

        if deferred:
            self._exchange.defer(self)
        elif self.training:
            self._ema_update(stats)
        if quantize.dtype != input.dtype:
            quantize = quantize.to(input.dtype)
        return quantize, scalars[0], ind.view(*input.shape[:-1]), scalars[1]

    def reduce_ema_stats(self, stats: torch.Tensor) -> torch.Tensor:
        """Sum the packed ``[counts (K) | embed_sum (K*D, code-major)]`` buffer over the
        ranks of ``stats_process_group`` in place (one collective per quantiser per step;
        NCCL on GPUs).  A no-op outside torch.distributed or with ``sync_ema_stats`` off."""
        import torch.distributed as dist
        if (self.sync_ema_stats and dist.is_available() and dist.is_initialized()
                and dist.get_world_size(self.stats_process_group) > 1):
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.stats_process_group)
        return stats

    def _ema_update(self, stats: torch.Tensor) -> None:
        """bottleneck.py:79-92.  With torch.distributed initialised the packed
        ``[counts | embed_sum]`` buffer is summed over ranks first, so N ranks update like
        one process on the concatenated batch (the reference itself never
        reduces these and lets DDP broadcast rank 0's buffers instead)."""
        self.reduce_ema_stats(stats)
        self._apply_ema(stats)

    def _apply_ema(self, stats: torch.Tensor) -> None:
        """``isi_vq_ema_update`` on (already reduced) packed statistics."""
        for buf in (self.cluster_size, self.embed_avg, self.embed):
            if not buf.is_contiguous():
                raise RuntimeError("codebook buffers must be contiguous")
        _lib.invoke("isi_vq_ema_update", 
            stats.data_ptr(), self.cluster_size.data_ptr(), self.embed_avg.data_ptr(),
            self.embed.data_ptr(), self.dim, self.n_embed, float(self.decay), float(self.eps),
            _lib.stream_ptr(stats.device))
        self._cache.invalidate()


class UnquantizedBottleneck(QuantizedBottleneck):
    """``bottleneck.py:107-119`` of the reference (selected by ``disable_quantization``,
    vqvae.py:159-160): the features pass through unquantised; no indices, infinite perplexity,
    zero commitment term.  Keeps the codebook buffers so that checkpoints interchange."""

    def forward(self, input: torch.Tensor):
        diff = torch.zeros((1,), dtype=input.dtype, device=input.device)
        perplexity = torch.as_tensor([math.inf], device=input.device)
        return input, diff, None, perplexity

    def assign(self, input: torch.Tensor):
        raise NotImplementedError("an unquantised bottleneck has no codes")

    def embed_code(self, embed_ind):
        raise NotImplementedError
