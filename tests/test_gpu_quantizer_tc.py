"""GPU parity of the tcgen05 (3xTF32) nearest-code kernel, called explicitly."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
from oracle import quantizer_oracle as qo
from test_gpu_quantizer import assert_indices_match, make

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n_embed,rows", [(512, 4096), (512, 40960), (512, 4097), (512, 100000),
                                          (64, 8192), (1000, 5000), (4096, 8192)])
def test_tcgen05_matches_fp64_outside_near_ties(n_embed, rows):
    embed = synthetic.synthetic_codebook(64, n_embed)
    x = synthetic.synthetic_features(rows, embed)
    m = make(64, n_embed, embed, "tcgen05").eval()
    ind = m.assign(x.to(DEV))
    near, flipped = assert_indices_match(ind, x, embed)
    simt = make(64, n_embed, embed, "simt").eval().assign(x.to(DEV))
    agree = (simt == ind).float().mean().item()
    print(f"[tcgen05] K={n_embed} N={rows}: {near} near ties, {flipped} flipped vs FP64, "
          f"agreement with the FP32 SIMT kernel {agree:.6f}")
    assert agree > 0.999


def test_tcgen05_nchw_strided_input_and_scales():
    embed = synthetic.synthetic_codebook(64, 512) * 3.0
    x = synthetic.synthetic_features(8 * 64 * 8 * 4, embed).view(8, 64, 32, 64)   # [B,H,W,D]
    nchw = x.permute(0, 3, 1, 2).contiguous().to(DEV)                           # conv output
    view = nchw.permute(0, 2, 3, 1)
    assert not view.is_contiguous()
    m = make(64, 512, embed, "tcgen05").eval()
    ind = m.assign(view)
    assert_indices_match(ind, x.reshape(-1, 64), embed)
    # tiny and huge magnitudes keep FP32-equivalent accuracy (no FP16-style range limits)
    for scale in (1e-3, 1e3):
        m2 = make(64, 512, embed * scale, "tcgen05").eval()
        assert_indices_match(m2.assign((x * scale).to(DEV)), x.reshape(-1, 64) * scale, embed * scale)


def test_tcgen05_exact_ties_pick_lowest_index():
    embed = synthetic.synthetic_codebook(64, 512)
    embed[:, 300] = embed[:, 17]
    embed[:, 511] = embed[:, 0]
    x = synthetic.synthetic_features(8192, embed)
    x[5] = embed[:, 17]
    x[6] = embed[:, 511]
    m = make(64, 512, embed, "tcgen05").eval()
    ind = m.assign(x.to(DEV)).cpu()
    assert ind[5] == 17 and ind[6] == 0
    assert not ((ind == 300) | (ind == 511)).any()


def test_tcgen05_forward_end_to_end_and_rerun_is_deterministic():
    embed = synthetic.synthetic_codebook(64, 512)
    x = synthetic.synthetic_features(65536, embed).to(DEV)
    m = make(64, 512, embed, "tcgen05").eval()
    q1, d1, i1, p1 = m(x)
    q2, d2, i2, p2 = m(x)
    assert torch.equal(i1, i2) and torch.equal(q1, q2) and d1.item() == d2.item()
    assert torch.equal(q1, m.embed_code(i1)) or (q1 - m.embed_code(i1)).abs().max() < 1e-6


# ---- the CTA-pair (cta_group::2) kernel with the resident codebook -------------------------
@pytest.mark.parametrize("n_embed,rows", [(512, 4096), (512, 40960), (512, 4097), (512, 100003),
                                          (64, 8192), (300, 5000), (256, 9999)])
def test_pair_kernel_matches_fp64_outside_near_ties(n_embed, rows):
    embed = synthetic.synthetic_codebook(64, n_embed)
    x = synthetic.synthetic_features(rows, embed)
    m = make(64, n_embed, embed, "tcgen05_pair").eval()
    ind = m.assign(x.to(DEV))
    near, flipped = assert_indices_match(ind, x, embed)
    simt = make(64, n_embed, embed, "simt").eval().assign(x.to(DEV))
    agree = (simt == ind).float().mean().item()
    print(f"[pair] K={n_embed} N={rows}: {near} near ties, {flipped} flipped vs FP64, "
          f"agreement with the FP32 SIMT kernel {agree:.6f}")
    assert agree > 0.999


def test_pair_kernel_nchw_ties_and_determinism():
    embed = synthetic.synthetic_codebook(64, 512)
    embed[:, 300] = embed[:, 17]
    embed[:, 511] = embed[:, 0]
    x = synthetic.synthetic_features(8 * 64 * 8 * 4, embed).view(8, 64, 32, 64)
    x[0, 0, 5] = embed[:, 17]
    x[0, 0, 6] = embed[:, 511]
    view = x.permute(0, 3, 1, 2).contiguous().to(DEV).permute(0, 2, 3, 1)
    m = make(64, 512, embed, "tcgen05_pair").eval()
    ind = m.assign(view)
    assert_indices_match(ind, x.reshape(-1, 64), embed)
    flat = ind.reshape(-1).cpu()
    assert flat[5] == 17 and flat[6] == 0 and not ((flat == 300) | (flat == 511)).any()
    assert torch.equal(m.assign(view), ind)
    assert torch.equal(make(64, 512, embed, "tcgen05").eval().assign(view), ind)


# ---- the streaming CTA-pair kernel: codebooks that do not fit in shared memory --------------
@pytest.mark.parametrize("dim,n_embed,rows", [(64, 1000, 5000), (64, 4096, 8192), (128, 4096, 8192),
                                              (128, 512, 4097), (128, 200, 20000), (64, 640, 33333)])
def test_pair_stream_kernel_matches_fp64_outside_near_ties(dim, n_embed, rows):
    embed = synthetic.synthetic_codebook(dim, n_embed)
    x = synthetic.synthetic_features(rows, embed)
    m = make(dim, n_embed, embed, "tcgen05_pair_stream").eval()
    ind = m.assign(x.to(DEV))
    near, flipped = assert_indices_match(ind, x, embed)
    simt = make(dim, n_embed, embed, "simt").eval().assign(x.to(DEV))
    agree = (simt == ind).float().mean().item()
    print(f"[pair-stream] D={dim} K={n_embed} N={rows}: {near} near ties, {flipped} flipped vs FP64, "
          f"agreement with the FP32 SIMT kernel {agree:.6f}")
    assert agree > 0.999


def test_pair_stream_kernel_nchw_view_d128():
    embed = synthetic.synthetic_codebook(128, 1024)
    x = synthetic.synthetic_features(4 * 48 * 24, embed).view(4, 48, 24, 128)
    view = x.permute(0, 3, 1, 2).contiguous().to(DEV).permute(0, 2, 3, 1)
    m = make(128, 1024, embed, "tcgen05_pair_stream").eval()
    ind = m.assign(view)
    assert_indices_match(ind, x.reshape(-1, 128), embed)
    assert torch.equal(m.assign(view), ind)
