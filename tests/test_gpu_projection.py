"""GPU parity of the pre-quantiser projection kernel (isi_vq_project: concat + 1x1 conv + bias on
tcgen05 with the 3xTF32 split) against torch in FP64, and of the VQVAE wiring that uses it."""
import pytest
import torch

from oracle import parity
from torch import nn

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae import vqvae as vq

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5          # relative to the output's max-abs; 3xTF32 leaves ~1e-6


def _cl(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)


def _reference(conv, sources):
    x = torch.cat([s.double() for s in sources], 1)
    y = torch.nn.functional.conv2d(x, conv.weight.double(), conv.bias.double())
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("shape,c0,c1", [((4, 32, 4), 128, 0), ((3, 64, 8), 64, 128), ((5, 37, 11), 64, 128),
                                         ((444, 64, 8), 64, 128), ((444, 32, 4), 128, 0), ((2, 50, 9), 192, 64)])
def test_projection_matches_fp64(shape, c0, c1):
    b, h, w = shape
    torch.manual_seed(1)
    conv = nn.Conv2d(c0 + c1, 64, 1).to(DEV)
    sources = [_cl(b, c0, h, w, seed=2)] + ([_cl(b, c1, h, w, seed=3)] if c1 else [])
    proj = vq.PointwiseProjection(conv)
    proj.min_rows = 1
    with torch.no_grad():
        assert proj.usable(sources)
        got = proj(sources)
        want = _reference(conv, sources)
    assert got.shape == (b, h, w, 64) and got.is_contiguous()
    err = (got.double() - want).abs().max() / want.abs().max()
    assert err <= TOL, float(err)


def test_strided_rows_folded_bias_and_weight_updates():
    torch.manual_seed(4)
    conv = nn.Conv2d(192, 64, 1).to(DEV)
    wide = _cl(3, 256, 40, 8, seed=5)
    a, b = wide[:, :64], wide[:, 128:256]               # row stride 256, channel offsets 0 and 128
    proj = vq.PointwiseProjection(conv)
    proj.min_rows = 1
    folded = torch.randn(64, device=DEV)
    with torch.no_grad():
        assert proj.usable([a, b])
        got = proj([a, b], folded_bias=folded)
        want = _reference(conv, [a + folded.view(1, -1, 1, 1), b])
        assert (got.double() - want).abs().max() <= TOL * want.abs().max()
        conv.weight.mul_(0.5)                             # in-place update: the prepared image is rebuilt
        conv.bias.add_(1.0)
        got = proj([a, b])
        want = _reference(conv, [a, b])
        assert (got.double() - want).abs().max() <= TOL * want.abs().max()
        # what does not qualify stays with the stock modules
        assert not proj.usable([a.contiguous(), b])       # NCHW-contiguous source
        assert not proj.usable([wide[:, :32], wide[:, 32:192]])
    with torch.enable_grad():
        assert not proj.usable([a, b])


def test_encode_codes_with_and_without_the_projection_kernel():
    """channels_last extraction uses the kernel (and folds dec_t's last bias into it); the
    stock-module path gives the same codes outside near ties."""
    torch.manual_seed(0)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2},
                     adapt_quantized_durations=False).to(DEV).eval().to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(9)
    spec = torch.randn(16, 2, 1024, 128, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    tf32, torch.backends.cudnn.allow_tf32 = torch.backends.cudnn.allow_tf32, False
    try:
        with torch.no_grad():
            calls = vq._lib.launch_counts["isi_vq_project"]
            id_t, id_b = model.encode_codes(spec)
            assert vq._lib.launch_counts["isi_vq_project"] == calls + 2
            full = model.encode(spec)
            assert torch.equal(full[3], id_t) and torch.equal(full[4], id_b)
            fused = parity.encode_with_features(model, spec)
            vq.fused_inference = False
            try:
                stock = parity.encode_with_features(model, spec)
                ref_t, ref_b = model.encode_codes(spec)
            finally:
                vq.fused_inference = True
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    # every code that differs from the stock-module path is a near tie on the stock features
    rep_t, rep_b, n_b = parity.explain_code_maps(stock, fused, id_t, id_b, model.quantize_t.embed.cpu(),
                                                 model.quantize_b.embed.cpu())
    print(f"[projection vs stock modules] top: {rep_t}; bottom ({n_b}/16 notes): {rep_b}")
    assert torch.equal(ref_t, stock[1]) and torch.equal(ref_b, stock[3])
    assert rep_t.unexplained == 0 and rep_b.unexplained == 0 and n_b >= 8


@pytest.mark.parametrize("factors", [{"bottom": 8, "top": 4}, {"bottom": 4, "top": 2}])
def test_other_resolution_configs_use_the_kernel(factors):
    """The in-tree configurations besides the deployed one: larger code maps,
    same 128 / 64 + 128 channel projections."""
    torch.manual_seed(0)
    model = vq.VQVAE(in_channel=2, resolution_factors=factors,
                     adapt_quantized_durations=False).to(DEV).eval().to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(3)
    spec = torch.randn(8, 2, 256, 64, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    tf32, torch.backends.cudnn.allow_tf32 = torch.backends.cudnn.allow_tf32, False
    try:
        with torch.no_grad():
            calls = vq._lib.launch_counts["isi_vq_project"]
            id_t, id_b = model.encode_codes(spec)
            assert vq._lib.launch_counts["isi_vq_project"] == calls + 2
            fused = parity.encode_with_features(model, spec)
            vq.fused_inference = False
            try:
                stock = parity.encode_with_features(model, spec)
                ref_t, ref_b = model.encode_codes(spec)
            finally:
                vq.fused_inference = True
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert id_t.shape == ref_t.shape and id_b.shape == ref_b.shape
    rep_t, rep_b, n_b = parity.explain_code_maps(stock, fused, id_t, id_b, model.quantize_t.embed.cpu(),
                                                 model.quantize_b.embed.cpu())
    print(f"[projection, {factors}] top: {rep_t}; bottom ({n_b}/8 notes): {rep_b}")
    assert rep_t.unexplained == 0 and rep_b.unexplained == 0 and n_b >= 4
