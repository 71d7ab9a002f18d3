"""Stress / determinism test of the EMA statistics kernels (``isi_vq_gather_stats``, the segmented
reduction behind bottleneck.py:79-85), instead of arguing with racecheck: the role-specialised
kernel hands rows from a producer warp to a dispatcher warp to 16 consumer warps through
mbarriers, and a lost or doubled hand-off would show as a wrong COUNT -- counts are integers held
in FP32, so they must be bit-identical over hundreds of repeats and equal to ``bincount``.

200 repeats x {skewed, uniform, one code, two codes 50/50} x n_rows in {63, 64, 65, 4097, 2^20+1}:
  * counts bit-identical every time and equal to the exact histogram;
  * embed_sum within 1e-6 of an FP64 reduction (relative to the largest sum);
  * the commitment term (diff) bit-identical every time (fixed-order reduction);
  * the dequantised rows exactly x + (E^T[ind] - x), the reference's straight-through value."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200 import _lib
from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
DIM, K = 64, 512


def _codes(kind: str, n: int, gen: torch.Generator) -> torch.Tensor:
    if kind == "uniform":
        return torch.randint(0, K, (n,), generator=gen)
    if kind == "one":
        return torch.full((n,), 137, dtype=torch.int64)
    if kind == "two":
        return torch.where(torch.arange(n) % 2 == 0, 3, 400).to(torch.int64)
    # skewed: a quarter of the rows on one code (silence), the rest Zipf-like
    ranks = torch.arange(1, K + 1, dtype=torch.float64)
    probs = 1.0 / ranks
    probs[0] = probs.sum() / 3.0
    return torch.multinomial((probs / probs.sum()).float(), n, replacement=True, generator=gen)


def _gather_stats(module: QuantizedBottleneck, x: torch.Tensor, ind: torch.Tensor):
    """The training-mode call of QuantizedBottleneck._forward_impl, with the codes given."""
    lib = _lib.load()
    n_rows = x.shape[0]
    prepared = module._cache.get(module.embed)
    layout = _lib.rows_layout(x)
    quantize = torch.empty_like(x)
    stats = torch.zeros(K * (1 + DIM), dtype=torch.float32, device=DEV)
    ws_bytes = lib.isi_vq_gather_workspace_bytes(n_rows, DIM)
    workspace = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=DEV)
    stream = _lib.stream_ptr(DEV)
    _lib.invoke("isi_vq_gather_stats", x.data_ptr(), layout, ind.data_ptr(), n_rows, DIM, K,
                prepared.data_ptr(), quantize.data_ptr(), _lib.rows_layout(quantize), stats.data_ptr(), 0,
                workspace.data_ptr(), workspace.numel() * 8, None, stream)
    scalars = torch.empty(2, dtype=torch.float32, device=DEV)
    _lib.invoke("isi_vq_finish", workspace.data_ptr(), n_rows, DIM, K, stats.data_ptr(), scalars.data_ptr(),
                scalars.data_ptr() + 4, stream)
    return quantize, stats, scalars


@pytest.mark.parametrize("kind", ["skewed", "uniform", "one", "two"])
@pytest.mark.parametrize("n_rows", [63, 64, 65, 4097, (1 << 20) + 1])
def test_statistics_are_exact_and_repeatable(kind, n_rows):
    gen = torch.Generator().manual_seed(1000 + n_rows % 977)
    embed = synthetic.synthetic_codebook(DIM, K)
    module = QuantizedBottleneck(DIM, K).to(DEV).train()
    module.embed.copy_(embed)
    ind_cpu = _codes(kind, n_rows, gen)
    x_cpu = torch.randn(n_rows, DIM, generator=gen) * 0.7 + embed.t()[ind_cpu] * 0.5
    x, ind = x_cpu.to(DEV), ind_cpu.to(DEV)
    want_counts = torch.bincount(ind_cpu, minlength=K).float()
    want_sum = torch.zeros(K, DIM, dtype=torch.float64).index_add_(0, ind_cpu, x_cpu.double())
    repeats = 200 if n_rows < 100000 else 60
    first = None
    for rep in range(repeats):
        quantize, stats, scalars = _gather_stats(module, x, ind)
        counts, embed_sum = stats[:K], stats[K:].view(K, DIM)
        if first is None:
            first = (counts.clone(), embed_sum.clone(), scalars.clone())
            assert torch.equal(counts.cpu(), want_counts), "counts differ from the exact histogram"
            err = (embed_sum.cpu().double() - want_sum).abs().max() / want_sum.abs().max()
            assert err <= 1e-6, f"embed_sum relative error {err:.3g}"
            # the straight-through value x + (q - x) of bottleneck.py:95, rounding included
            assert torch.equal(quantize.cpu(), x_cpu + (embed.t()[ind_cpu] - x_cpu)), "dequantised rows"
            want_diff = ((embed.t()[ind_cpu].double() - x_cpu.double()) ** 2).mean()
            assert abs(scalars[0].item() - want_diff.item()) <= 1e-5 * want_diff.item()
        else:
            assert torch.equal(counts, first[0]), f"counts changed on repeat {rep}"
            assert torch.equal(scalars, first[2]), f"diff / perplexity changed on repeat {rep}"
            # the sums go through floating-point atomics (one per CTA and used code): order may
            # vary between launches, so they are compared with the FP64 reduction, not bitwise
            err = (embed_sum.cpu().double() - want_sum).abs().max() / want_sum.abs().max()
            assert err <= 1e-6, f"embed_sum relative error {err:.3g} on repeat {rep}"
    torch.cuda.synchronize()


def test_nan_rows_get_a_valid_code():
    """a NaN / Inf row never wins a comparison in the search; it must still come
    back with an index inside [0, K) (the reference's ``(-dist).max(1)`` always does)."""
    embed = synthetic.synthetic_codebook(DIM, K)
    for algo, rows in (("simt", 130), ("auto", 130), ("auto", 8192)):
        module = QuantizedBottleneck(DIM, K).to(DEV).eval()
        module.embed.copy_(embed)
        module.assign_algo = algo
        x = synthetic.synthetic_features(rows, embed).to(DEV)
        x[5] = float("nan")
        x[17, 3] = float("inf")
        ind = module.assign(x)
        assert int(ind.min()) >= 0 and int(ind.max()) < K, (algo, rows, int(ind.min()), int(ind.max()))
        quant, diff, ind2, perp = module(x)
        assert int(ind2.min()) >= 0 and int(ind2.max()) < K
