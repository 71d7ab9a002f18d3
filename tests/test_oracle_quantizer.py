"""The quantiser oracle against (a) the committed fixtures generated from the
unmodified reference class and (b) the live reference when the checkout exists."""
import warnings

import numpy as np
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from oracle import quantizer_oracle as qo
from oracle import ref_loader


def _state(embed):
    e = torch.as_tensor(embed).clone()
    return qo.CodebookState(e, torch.zeros(e.shape[1]), e.clone())


def test_eval_fixture_cfg1(golden_dir):
    g = np.load(golden_dir / "quantizer_eval_cfg1.npz")
    st = _state(g["embed"])
    for name in ("top", "bottom"):
        x = torch.from_numpy(g[f"x_{name}"])
        quant, diff, ind, perp = qo.forward(st, x, training=False)
        assert ind.dtype == torch.int64 and ind.shape == x.shape[:-1]
        np.testing.assert_array_equal(ind.numpy(), g[f"ind_{name}"])
        np.testing.assert_array_equal(quant.numpy(), g[f"quantize_{name}"])
        np.testing.assert_allclose(diff.numpy(), g[f"diff_{name}"], rtol=1e-6)
        np.testing.assert_allclose(perp.numpy(), g[f"perplexity_{name}"], rtol=1e-6)


def test_train_fixture_three_steps(golden_dir):
    g = np.load(golden_dir / "quantizer_train_3steps.npz")
    st = _state(g["embed0"])
    for step in range(3):
        x = torch.from_numpy(g[f"x{step}"])
        quant, diff, ind, perp = qo.forward(st, x, training=True)
        np.testing.assert_array_equal(ind.numpy(), g[f"ind{step}"])
        np.testing.assert_array_equal(quant.numpy(), g[f"quantize{step}"])
        np.testing.assert_allclose(st.cluster_size.numpy(), g[f"cluster_size_after{step}"], rtol=1e-6)
        np.testing.assert_allclose(st.embed_avg.numpy(), g[f"embed_avg_after{step}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(st.embed.numpy(), g[f"embed_after{step}"], rtol=1e-5, atol=1e-7)


def test_edge_fixture_ties_and_strided_input(golden_dir):
    g = np.load(golden_dir / "quantizer_edge.npz")
    st = _state(g["embed"])
    x = torch.from_numpy(g["nchw"]).permute(0, 2, 3, 1)
    assert not x.is_contiguous()
    quant, diff, ind, perp = qo.forward(st, x)
    np.testing.assert_array_equal(ind.numpy(), g["ind"])
    assert ind[0, 0, 0] == 3 and ind[1, 2, 2] == 0      # duplicates: lowest index wins
    np.testing.assert_array_equal(quant.numpy(), g["quantize"])
    np.testing.assert_allclose(diff.numpy(), g["diff"], rtol=1e-6)
    np.testing.assert_array_equal(
        qo.dequantise(torch.from_numpy(g["codes"]), st.embed).numpy(), g["looked_up"])


def test_fp64_gap_flags_only_a_few_positions():
    embed = synthetic.synthetic_codebook()
    x = synthetic.synthetic_features(8192, embed)
    ind32 = qo.assign(x, embed)
    ind64, gap = qo.assign_fp64(x, embed)
    clear = gap > 1e-5
    assert (ind32[clear] == ind64[clear]).all()
    assert (~clear).float().mean() < 5e-3


def test_corruption_wraps_and_is_pm_one():
    ind = torch.tensor([0, 1, 510, 511])
    g = torch.Generator().manual_seed(0)
    for _ in range(20):
        out = qo.corrupt(ind, 512, [0.1, 0.8, 0.1], g)
        d = (out - ind) % 512
        assert set(d.tolist()) <= {0, 1, 511}


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("dim,n_embed,rows", [(64, 512, 640), (128, 4096, 256), (8, 20, 37)])
def test_live_reference_eval_and_train(dim, n_embed, rows):
    ref = ref_loader.load_reference_bottleneck()
    embed = synthetic.synthetic_codebook(dim, n_embed)
    x = synthetic.synthetic_features(rows, embed, 42)
    m = ref.QuantizedBottleneck(dim, n_embed)
    m.embed.copy_(embed)
    m.embed_avg.copy_(embed)
    st = _state(embed)
    for training in (False, True, True):
        m.train(training)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rq, rd, ri, rp = m(x)
        oq, od, oi, op = qo.forward(st, x, training=training)
        assert torch.equal(ri, oi)
        assert torch.equal(rq, oq)
        torch.testing.assert_close(rd, od, rtol=1e-6, atol=0)
        torch.testing.assert_close(rp, op, rtol=1e-6, atol=0)
        torch.testing.assert_close(m.embed, st.embed, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(m.cluster_size, st.cluster_size, rtol=1e-6, atol=0)
        torch.testing.assert_close(m.embed_avg, st.embed_avg, rtol=1e-6, atol=1e-7)
        x = synthetic.synthetic_features(rows, st.embed, 43)
