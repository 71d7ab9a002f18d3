"""Seeded synthetic workloads: NSynth-shaped notes, pre-quantiser
features, codebooks and codemaps.  Generated on the CPU with ``torch.Generator``
so the same bytes reach the CUDA path, the oracle and the golden fixtures."""
import math

import torch

AUDIO_SEED = 20200117      # the reference's own reproducible seed (create_nsynth_dataset_split.py:12)
FEATURE_SEED_EVAL = 1234
FEATURE_SEED_TRAIN = 4321
CODEBOOK_SEED = 0


def synthetic_notes(batch: int, n_samples: int = 64000, fs_hz: int = 16000,
                    seed: int = 20200117) -> torch.Tensor:
    """Seeded NSynth-shaped notes: 8 decaying harmonics of a MIDI
    pitch in 24..84 plus a -60 dB noise floor, peak-normalised to 0.9."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / fs_hz
    pitch = torch.randint(24, 85, (batch,), generator=g).double()
    f0 = 440.0 * torch.pow(torch.tensor(2.0, dtype=torch.float64), (pitch - 69.0) / 12.0)
    tau = 0.2 + 1.8 * torch.rand(batch, generator=g, dtype=torch.float64)
    amp = torch.rand(batch, 8, generator=g, dtype=torch.float64) / torch.arange(1, 9).double()
    phi = 2 * math.pi * torch.rand(batch, 8, generator=g, dtype=torch.float64)
    out = torch.zeros(batch, n_samples, dtype=torch.float64)
    for h in range(8):
        fh = f0 * (h + 1)
        keep = (fh < fs_hz / 2).double()
        out += (keep * amp[:, h])[:, None] * torch.sin(
            2 * math.pi * fh[:, None] * t[None, :] + phi[:, h:h + 1])
    out *= torch.exp(-t[None, :] / tau[:, None])
    out += 1e-3 * torch.randn(batch, n_samples, generator=g, dtype=torch.float64)
    out *= 0.9 / out.abs().amax(dim=1, keepdim=True)
    return out.float()


def synthetic_codebook(dim: int = 64, n_embed: int = 512, seed: int = CODEBOOK_SEED,
                       variance: float = 1.0) -> torch.Tensor:
    """``[D, K]`` like the reference initialiser (bottleneck.py:46-48)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(dim, n_embed, generator=g) * math.sqrt(variance)


def synthetic_features(n_rows: int, embed: torch.Tensor, seed: int = FEATURE_SEED_EVAL
                       ) -> torch.Tensor:
    """``[n_rows, D]`` pre-quantiser rows: a quarter N(0,1), a quarter 0.1 N(0,1), half
    near-codeword rows ``E^T[randint] + 0.05 N(0,1)`` (the near-tie probe mix)."""
    dim, n_embed = embed.shape
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, dim, generator=g)
    q = n_rows // 4
    x[q:2 * q] *= 0.1
    pick = torch.randint(0, n_embed, (n_rows - 2 * q,), generator=g)
    x[2 * q:] = embed.t()[pick] + 0.05 * x[2 * q:]
    return x[torch.randperm(n_rows, generator=g)].contiguous()


def synthetic_codemaps(batch: int, n_embed: int = 512, top=(32, 4), bottom=(64, 8),
                       seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, n_embed, (batch, *top), generator=g),
            torch.randint(0, n_embed, (batch, *bottom), generator=g))
