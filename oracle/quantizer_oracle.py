"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference EMA quantiser.

Restates ``QuantizedBottleneck`` of the reference
(``interactive_spectrogram_inpainting/vqvae/bottleneck.py:30-104``) as plain
functions over explicit state, in FP32 torch-CPU arithmetic (the reference is
torch too, so the same BLAS does the contraction) plus an FP64 variant used to
measure near-tie gaps.  Nothing under ``interactive_spectrogram_inpainting_b200``
imports this file; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / the timed CPU baseline.

Parity status: PINNED.  ``tests/test_oracle_quantizer.py`` checks every function
here against (a) the unmodified reference class imported from the read-only
checkout when it is present and (b) the committed fixtures in
``tests/golden/quantizer_*.npz`` that ``oracle/make_golden.py`` generated from
that class.
"""
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass
class CodebookState:
    """The three buffers of bottleneck.py:49-51 (shapes [D,K], [K], [D,K])."""
    embed: torch.Tensor
    cluster_size: torch.Tensor
    embed_avg: torch.Tensor

    @staticmethod
    def fresh(dim: int, n_embed: int, seed: int = 0,
              initial_variance: float = 1.0) -> "CodebookState":
        # bottleneck.py:46-51: randn(dim, n_embed) * sqrt(var); zeros; clone
        g = torch.Generator().manual_seed(seed)
        e = torch.randn(dim, n_embed, generator=g) * float(np.sqrt(initial_variance))
        return CodebookState(e, torch.zeros(n_embed), e.clone())

    def clone(self) -> "CodebookState":
        return CodebookState(self.embed.clone(), self.cluster_size.clone(),
                             self.embed_avg.clone())


def distances(rows: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:56-60 -- ||x||^2 - 2 x.E + ||E||^2, same association order."""
    sq_rows = (rows * rows).sum(dim=1, keepdim=True)
    cross = (2 * rows) @ embed
    sq_codes = (embed * embed).sum(dim=0, keepdim=True)
    return sq_rows - cross + sq_codes


def assign(rows: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:61 -- index of the max of -dist (first occurrence on ties)."""
    return torch.max(-distances(rows, embed), dim=1).indices


def assign_fp64(rows: torch.Tensor, embed: torch.Tensor
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """FP64 nearest code and the relative gap (d2 - d1) / |d1| to the runner-up.

    Used by the parity tests to decide which positions are "near ties"
    (BASELINE.json north_star: relative gap <= 1e-5) where an FP32 evaluation
    order may legitimately pick the other code.
    """
    d = distances(rows.double(), embed.double())
    two = torch.topk(d, k=2, dim=1, largest=False)
    d1, d2 = two.values[:, 0], two.values[:, 1]
    gap = (d2 - d1) / d1.abs().clamp_min(1e-300)
    return two.indices[:, 0], gap


def corrupt(indices: torch.Tensor, n_embed: int, weights: Sequence[float],
            generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """bottleneck.py:63-73 -- add U{-1,0,+1} drawn with ``weights`` then wrap."""
    draw = torch.multinomial(torch.tensor(list(weights), dtype=torch.float32),
                             indices.numel(), replacement=True,
                             generator=generator)
    shift = (draw - 1).reshape(indices.shape).to(indices.device)
    return (indices + shift) % n_embed


def dequantise(indices: torch.Tensor, embed: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:103-104 -- rows of embed^T selected by ``indices``."""
    return embed.t()[indices]


def ema_update(state: CodebookState, rows: torch.Tensor, indices: torch.Tensor,
               decay: float, eps: float,
               counts: Optional[torch.Tensor] = None,
               embed_sum: Optional[torch.Tensor] = None) -> None:
    """bottleneck.py:79-92, in place on ``state``.

    ``counts``/``embed_sum`` may be supplied (e.g. summed over ranks) to model
    the all-reduced multi-GPU update of SURVEY.md F3; by default they are the
    one-hot column sums (bottleneck.py:81) and x^T.onehot (bottleneck.py:83).
    """
    n_embed = state.embed.shape[1]
    if counts is None or embed_sum is None:
        onehot = torch.nn.functional.one_hot(indices.reshape(-1), n_embed).to(rows.dtype)
        counts = onehot.sum(0)
        embed_sum = rows.t() @ onehot
    state.cluster_size.mul_(decay).add_(counts, alpha=1 - decay)      # :80-82
    state.embed_avg.mul_(decay).add_(embed_sum, alpha=1 - decay)      # :84-85
    total = state.cluster_size.sum()                                  # :86
    smoothed = (state.cluster_size + eps) / (total + n_embed * eps) * total  # :87-90
    state.embed.copy_(state.embed_avg / smoothed.unsqueeze(0))        # :91-92


def perplexity(indices: torch.Tensor, n_embed: int) -> torch.Tensor:
    """bottleneck.py:97-100 -- exp(-sum p log max(p, 1e-7)), p = usage frequency."""
    p = torch.bincount(indices.reshape(-1), minlength=n_embed).to(torch.float32)
    p = p / indices.numel()
    return torch.exp(-(p * torch.log(p.clamp(min=1e-7))).sum())


def forward(state: CodebookState, x: torch.Tensor, training: bool = False,
            decay: float = 0.99, eps: float = 1e-5,
            corruption_weights: Optional[Sequence[float]] = None,
            generator: Optional[torch.Generator] = None):
    """bottleneck.py:53-101 end to end: (quantize, diff, embed_ind, perplexity).

    ``x`` is ``[..., D]`` with arbitrary strides.  The gather uses the codebook
    *before* the EMA update (bottleneck.py:77 precedes :79-92).
    """
    dim, n_embed = state.embed.shape
    rows = x.reshape(-1, dim)
    ind = assign(rows, state.embed)
    if training and corruption_weights is not None:
        ind = corrupt(ind, n_embed, corruption_weights, generator)
    q = dequantise(ind.view(*x.shape[:-1]), state.embed)
    if training:
        ema_update(state, rows, ind, decay, eps)
    diff = ((q - x) ** 2).mean()                                      # :94
    out = x + (q - x)                                                 # :95 (forward value)
    return out, diff, ind.view(*x.shape[:-1]), perplexity(ind, n_embed)


# --------------------------------------------------------------------------
# numpy twin of the integer/byte part (index bookkeeping), used by host tests
# --------------------------------------------------------------------------
def usage_histogram(indices: np.ndarray, n_embed: int) -> np.ndarray:
    return np.bincount(np.asarray(indices).reshape(-1), minlength=n_embed).astype(np.int64)


class OracleBottleneck(torch.nn.Module):
    """nn.Module face of the oracle with the reference's buffers, so the product's VQVAE
    wiring can be run on the CPU in tests and in bench.py's CPU-baseline legs."""

    def __init__(self, dim, n_embed, decay=0.99, eps=1e-5, embeddings_initial_variance=1,
                 corruption_weights=None):
        super().__init__()
        self.dim, self.n_embed, self.decay, self.eps = dim, n_embed, decay, eps
        self.corruption_weights = corruption_weights
        st = CodebookState.fresh(dim, n_embed, seed=torch.seed() % (2 ** 31),
                                 initial_variance=embeddings_initial_variance)
        self.register_buffer('embed', st.embed)
        self.register_buffer('cluster_size', st.cluster_size)
        self.register_buffer('embed_avg', st.embed_avg)

    def forward(self, input):
        st = CodebookState(self.embed, self.cluster_size, self.embed_avg)
        with torch.no_grad():
            out, diff, ind, perp = forward(st, input.detach(), self.training, self.decay,
                                           self.eps, self.corruption_weights)
        if input.requires_grad:
            q = out
            diff = ((q - input) ** 2).mean()
            out = input + (q - input).detach()
        return out, diff, ind, perp

    def embed_code(self, embed_id):
        return dequantise(embed_id, self.embed)
