"""CPU checks around the pre-quantiser projection (isi_vq_project): which tensors qualify for
the kernel, that everything else keeps the stock modules (same result as the reference wiring),
and the error of its 3xTF32 arithmetic, emulated with integer rounding of FP32 bit patterns."""
import numpy as np
import torch
from torch import nn

from interactive_spectrogram_inpainting_b200.vqvae import vqvae as vq


def test_rows_view_of_channels_last_tensors():
    x = torch.zeros(3, 128, 5, 7).contiguous(memory_format=torch.channels_last)
    assert vq._as_rows(x) == 128
    assert vq._as_rows(x[:, :64]) == 128                       # channel slice: same row stride
    assert vq._as_rows(torch.zeros(3, 128, 5, 7)) is None      # NCHW storage
    assert vq._as_rows(x[..., :5]) is None                     # cropped width: rows are not uniform
    assert vq._as_rows(torch.zeros(1, 64, 1, 9).contiguous(memory_format=torch.channels_last)) == 64


def test_cpu_and_grad_inputs_keep_the_stock_modules():
    proj = vq.PointwiseProjection(nn.Conv2d(192, 64, 1))
    a = torch.zeros(2, 64, 8, 8).contiguous(memory_format=torch.channels_last)
    b = torch.zeros(2, 128, 8, 8).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        assert not proj.usable([a, b])                         # CPU tensors: no kernel, no fallback inside it
    assert not vq.PointwiseProjection(nn.Conv2d(192, 32, 1)).usable([a, b])


def test_channels_last_model_on_cpu_matches_nchw_wiring():
    """The fast path must not change what the modules compute when it does not apply."""
    torch.manual_seed(0)
    from oracle import quantizer_oracle as qo
    model = vq.VQVAE(in_channel=2, resolution_factors={'bottom': 16, 'top': 2},
                     adapt_quantized_durations=False, bottleneck_cls=qo.OracleBottleneck).eval()
    spec = torch.randn(1, 2, 256, 64)
    with torch.no_grad():
        ref = model.encode(spec)
        model.to(memory_format=torch.channels_last)
        got = model.encode(spec.contiguous(memory_format=torch.channels_last))
    assert torch.equal(ref[3], got[3]) and torch.equal(ref[4], got[4])
    assert torch.allclose(ref[0], got[0], atol=1e-6) and torch.allclose(ref[1], got[1], atol=1e-6)


def _to_tf32(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits."""
    bits = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    bits = (bits + 0x1000) & 0xFFFFE000
    return bits.astype(np.uint32).view(np.float32)


def test_three_term_tf32_split_is_fp32_equivalent():
    """out = f_hi w_hi + f_hi w_lo + f_lo w_hi with exact products and FP32 accumulation (what
    the tensor cores do) differs from the FP64 result by ~1e-6 of the output's scale: the level
    of an FP32 GEMM, three orders below a single TF32 pass."""
    rng = np.random.default_rng(0)
    f = rng.standard_normal((512, 192)).astype(np.float32)
    w = (rng.standard_normal((64, 192)) / np.sqrt(192)).astype(np.float32)
    f_hi, w_hi = _to_tf32(f), _to_tf32(w)
    f_lo, w_lo = _to_tf32(f - f_hi), _to_tf32(w - w_hi)
    assert np.abs(f - f_hi - f_lo).max() <= 2.0 ** -21 * np.abs(f).max()      # the split loses < 2^-21
    acc = np.zeros((512, 64), dtype=np.float32)
    for a, b in ((f_lo, w_hi), (f_hi, w_lo), (f_hi, w_hi)):                    # small terms first
        acc += (a.astype(np.float64) @ b.astype(np.float64).T).astype(np.float32)
    exact = f.astype(np.float64) @ w.astype(np.float64).T
    scale = np.abs(exact).max()
    assert np.abs(acc - exact).max() <= 2e-6 * scale
    one_pass = f_hi.astype(np.float64) @ w_hi.astype(np.float64).T
    assert np.abs(one_pass - exact).max() > 1e-4 * scale
