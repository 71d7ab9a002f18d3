"""GPU checks of the extraction callers around the hot path: the conv stacks with cuDNN's
fused conv+bias(+skip)+ReLU against the stock module stacks, and ``encode_codes`` end to end."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae import vqvae as vq
from oracle import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture
def fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old
    vq.fused_inference = True


@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_conv_stacks_equal_stock_stacks(fp32_convs, channels_last):
    torch.manual_seed(5)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}).to(DEV).eval()
    x = torch.randn(3, 2, 256, 64, device=DEV)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        vq.fused_inference = False
        stock_b = model.enc_b(x)
        stock_t = model.enc_t(stock_b)
        stock_d = model.dec_t(stock_t[:, :64])
        vq.fused_inference = True
        fused_b = model.enc_b(x)
        fused_t = model.enc_t(stock_b)
        fused_d = model.dec_t(stock_t[:, :64])
    for got, want in ((fused_b, stock_b), (fused_t, stock_t), (fused_d, stock_d)):
        assert got.shape == want.shape
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    # training (grad enabled) keeps the stock modules: gradients flow
    y = model.enc_t(stock_b.requires_grad_())
    y.sum().backward()
    assert stock_b.grad is not None


def test_encode_codes_same_codes_fused_and_stock(fp32_convs, capsys):
    """Same code maps from both conv paths (they differ by FP32 rounding only): every differing
    code is explained by the FP64 distances on the stock path's features and the measured
    feature difference (oracle/parity.py); no agreement percentage."""
    torch.manual_seed(6)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}).to(DEV).eval()
    spec = torch.randn(4, 2, 1024, 128, device=DEV)
    with torch.no_grad():
        vq.fused_inference = False
        stock = parity.encode_with_features(model, spec)
        t0, b0 = model.encode_codes(spec)
        vq.fused_inference = True
        fused = parity.encode_with_features(model, spec)
        t1, b1 = model.encode_codes(spec)
    assert t0.shape == (4, 32, 4) and b0.shape == (4, 64, 8)
    assert torch.equal(t0, stock[1]) and torch.equal(b0, stock[3])
    rep_t, rep_b, n_b = parity.explain_code_maps(stock, fused, t1, b1, model.quantize_t.embed.cpu(),
                                                 model.quantize_b.embed.cpu())
    with capsys.disabled():
        print(f"\n[fused vs stock convs] top: {rep_t}\n[fused vs stock convs] bottom ({n_b}/4 notes): {rep_b}")
    assert rep_t.unexplained == 0 and rep_b.unexplained == 0 and n_b >= 2


def test_graphed_decode_code_equals_eager():
    torch.manual_seed(7)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}).to(DEV).eval()
    top, bottom = synthetic.synthetic_codemaps(2)
    top, bottom = top.to(DEV), bottom.to(DEV)
    graphed = vq.GraphedDecodeCode(model, top, bottom)
    with torch.no_grad():
        want = model.decode_code(top, bottom)
        torch.testing.assert_close(graphed(top, bottom), want, rtol=1e-3, atol=1e-4)
        top2 = (top + 17) % model.n_embed_t
        bottom2 = (bottom * 3 + 1) % model.n_embed_b
        want2 = model.decode_code(top2, bottom2)
        got2 = graphed(top2, bottom2).clone()
        torch.testing.assert_close(got2, want2, rtol=1e-3, atol=1e-4)
        assert not torch.allclose(got2, want)
    with pytest.raises(ValueError):
        graphed(top[:1], bottom[:1])
    with pytest.raises(RuntimeError):
        vq.GraphedDecodeCode(model.train(), top, bottom)


def test_space_to_depth_extraction_gives_the_same_codes(fp32_convs):
    """Front end writing 2x2 blocks + first conv as 3x3 stride 1 == the plain path."""
    from interactive_spectrogram_inpainting_b200 import extract
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    torch.manual_seed(8)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}).to(DEV).eval()
    model = model.to(memory_format=torch.channels_last)
    audio = synthetic.synthetic_notes(6)
    names = [f"n{i}" for i in range(6)]
    rows = {}
    for s2d in (False, True):
        helper = MelSpectrogramsHelper(channels_last=True, space_to_depth=s2d).to(DEV)
        loader = extract.SpectrogramBatches([(audio[:4].pin_memory(), names[:4]), (audio[4:].pin_memory(), names[4:])],
                                            helper, torch.device(DEV))
        rows[s2d] = extract.extract_codes(loader, model)
    assert [r.filename for r in rows[True]] == names
    # every code that differs between the two layouts is a near tie on the plain path's features
    with torch.no_grad():
        plain_spec = MelSpectrogramsHelper(channels_last=True).to(DEV).to_spectrogram(audio.to(DEV))
        plain = parity.encode_with_features(model, plain_spec)
        block_spec = MelSpectrogramsHelper.from_space_to_depth(
            MelSpectrogramsHelper(space_to_depth=True).to(DEV).to_spectrogram(audio.to(DEV)))
        blocks_pass = parity.encode_with_features(model, block_spec.contiguous(memory_format=torch.channels_last))
    import numpy as np
    for s2d in (False, True):
        got_t = torch.from_numpy(np.stack([r.top for r in rows[s2d]]))
        got_b = torch.from_numpy(np.stack([r.bottom for r in rows[s2d]]))
        rep_t, rep_b, n_b = parity.explain_code_maps(plain, blocks_pass, got_t, got_b, model.quantize_t.embed.cpu(),
                                                     model.quantize_b.embed.cpu())
        print(f"[space_to_depth={s2d}] top: {rep_t}; bottom ({n_b}/6 notes): {rep_b}")
        assert rep_t.unexplained == 0 and rep_b.unexplained == 0 and n_b >= 3
    with torch.no_grad():
        spec = MelSpectrogramsHelper(channels_last=True).to(DEV).to_spectrogram(audio.to(DEV))
        blocks = MelSpectrogramsHelper(space_to_depth=True).to(DEV).to_spectrogram(audio.to(DEV))
        torch.testing.assert_close(model.enc_b(blocks, space_to_depth=True), model.enc_b(spec),
                                   rtol=1e-4, atol=1e-5)


def test_code_extractor_graph_replay_equals_the_eager_pipeline():
    """``CodeExtractor`` (H2D one batch ahead, one CUDA-graph replay per batch, D2H) returns the
    rows of ``extract_codes`` over ``SpectrogramBatches``: same kernels, same order."""
    from interactive_spectrogram_inpainting_b200 import extract
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    torch.manual_seed(7)
    dev = torch.device(DEV)
    model = (vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)
             .to(dev).eval().to(memory_format=torch.channels_last))
    helper = MelSpectrogramsHelper(space_to_depth=True).to(dev)
    pcm = (synthetic.synthetic_notes(14) * 32767).round().to(torch.int16)
    batches = [(pcm[0:4].pin_memory(), [f"n{i}" for i in range(0, 4)]),
               (pcm[4:8].pin_memory(), [f"n{i}" for i in range(4, 8)]),
               (pcm[8:12].pin_memory(), [f"n{i}" for i in range(8, 12)]),
               (pcm[12:14].pin_memory(), [f"n{i}" for i in range(12, 14)])]      # ragged last batch
    want = extract.extract_codes(extract.SpectrogramBatches(batches, helper, dev), model)
    graphed = extract.CodeExtractor(helper, model, dev, cuda_graph=True)
    eager = extract.CodeExtractor(helper, model, dev, cuda_graph=False)
    streamed = []
    for run in range(2):                                     # second run replays the cached graphs
        got = graphed.run(batches, sink=streamed.extend)
        assert graphed.graph_failures == [], graphed.graph_failures
        assert [r.filename for r in got] == [r.filename for r in want]
        for g, w in zip(got, want):
            assert (g.top == w.top).all() and (g.bottom == w.bottom).all()
    assert len(streamed) == 2 * len(want)
    assert len(graphed._graphs) == 2 and all(graphed._graphs.values())     # one graph per batch shape
    for g, w in zip(eager.run(batches), want):
        assert (g.top == w.top).all() and (g.bottom == w.bottom).all()
    assert graphed.run([]) == []


def test_code_extractor_carries_attributes_and_names_through_the_prefetch_ring(fp32_convs):
    """extract_code.py:62-74: each row keeps ITS note's label-encoded attributes and file name.
    With uploads two batches ahead and a ragged last batch, a slot mix-up in the ring would pair a
    note's codes with another batch's attributes."""
    from interactive_spectrogram_inpainting_b200 import extract
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    torch.manual_seed(8)
    dev = torch.device(DEV)
    model = (vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)
             .to(dev).eval().to(memory_format=torch.channels_last))
    helper = MelSpectrogramsHelper(space_to_depth=True).to(dev)
    pcm = (synthetic.synthetic_notes(11) * 32767).round().to(torch.int16)
    pitch = torch.arange(11) + 40
    family = torch.arange(11) % 3
    cuts = [(0, 3), (3, 6), (6, 9), (9, 11)]
    batches = [(pcm[a:b].pin_memory(), [f"note_{i:02d}" for i in range(a, b)],
                {"pitch": pitch[a:b], "instrument_family_str": family[a:b]}) for a, b in cuts]
    alone = extract.CodeExtractor(helper, model, dev, prefetch_depth=1)
    for depth in (1, 2, 3):
        rows = extract.CodeExtractor(helper, model, dev, prefetch_depth=depth).run(batches)
        assert [r.filename for r in rows] == [f"note_{i:02d}" for i in range(11)]
        for i, row in enumerate(rows):
            assert set(row.attributes) == {"pitch", "instrument_family_str"}
            assert row.attributes["pitch"].shape == () and row.attributes["pitch"].dtype == torch.int64
            assert int(row.attributes["pitch"]) == 40 + i and int(row.attributes["instrument_family_str"]) == i % 3
            # the codes of note i, whichever batch and slot it travelled in
            single = alone.run([(pcm[i:i + 1].pin_memory(), [f"note_{i:02d}"])])[0]
            # (an identity check, not a parity check: cuDNN may pick another algorithm at batch 1,
            # so a near-tie may flip; another note's codes would differ almost everywhere)
            differing = int((row.top != single.top).sum()) + int((row.bottom != single.bottom).sum())
            assert differing <= 2, (depth, i, differing)


def test_transposed_plane_extraction_gives_the_same_codes(fp32_convs):
    """The extraction path of the bench: the front end writes the 2x2 blocks frequency-fastest,
    the encoder runs on the transposed plane with transposed filters (vqvae.TransposedFilters),
    the code maps come back as ``[B, F', T']``.  Same codes as the plain path, every difference a
    near tie on the plain path's features; the eager and the graphed extractor agree exactly."""
    from interactive_spectrogram_inpainting_b200 import extract
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    import numpy as np
    torch.manual_seed(9)
    model = vq.VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)
    model = model.to(DEV).eval().to(memory_format=torch.channels_last)
    audio = synthetic.synthetic_notes(6)
    pcm = (audio * 32767).round().to(torch.int16)
    names = [f"n{i}" for i in range(6)]
    batches = [(pcm[:4].pin_memory(), names[:4]), (pcm[4:].pin_memory(), names[4:])]
    helper_t = MelSpectrogramsHelper(space_to_depth="transposed").to(DEV)
    with torch.no_grad():
        # (1) the stacks alone: transposed plane == plain plane, transposed back
        spec = MelSpectrogramsHelper(channels_last=True).to(DEV).to_spectrogram(pcm.to(DEV))
        tblocks = helper_t.to_spectrogram(pcm.to(DEV))
        tf = model._transposed_filters
        enc_b = model.enc_b(spec)
        enc_b_t = model.enc_b(tblocks, space_to_depth=True, transposed=tf)
        torch.testing.assert_close(enc_b_t.transpose(2, 3), enc_b, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(model.enc_t(enc_b_t, transposed=tf).transpose(2, 3), model.enc_t(enc_b),
                                   rtol=1e-4, atol=1e-5)
        q = torch.randn(6, 64, 32, 4, device=DEV).contiguous(memory_format=torch.channels_last)
        torch.testing.assert_close(
            model.dec_t(q.transpose(2, 3).contiguous(memory_format=torch.channels_last), transposed=tf).transpose(2, 3),
            model.dec_t(q), rtol=1e-4, atol=1e-5)
        # (2) the code maps
        id_t, id_b = model.encode_codes(tblocks, space_to_depth="transposed")
        assert id_t.shape == (6, 32, 4) and id_b.shape == (6, 64, 8) and id_t.dtype == torch.int64
        plain = parity.encode_with_features(model, spec)
        rep_t, rep_b, n_b = parity.explain_code_maps(plain, plain, id_t.cpu(), id_b.cpu(),
                                                     model.quantize_t.embed.cpu(), model.quantize_b.embed.cpu())
        print(f"[transposed plane] top: {rep_t}; bottom ({n_b}/6 notes): {rep_b}")
        assert rep_t.unexplained == 0 and rep_b.unexplained == 0 and n_b >= 3
    # (3) through the extractors
    want_t, want_b = id_t.cpu().numpy(), id_b.cpu().numpy()
    for graph in (False, True):
        ex = extract.CodeExtractor(helper_t, model, torch.device(DEV), cuda_graph=graph)
        rows = ex.run(batches)
        assert ex.graph_failures == [] and [r.filename for r in rows] == names
        assert np.array_equal(np.stack([r.top for r in rows]), want_t)
        assert np.array_equal(np.stack([r.bottom for r in rows]), want_b)
    rows = extract.extract_codes(extract.SpectrogramBatches(batches, helper_t, torch.device(DEV)), model)
    assert np.array_equal(np.stack([r.top for r in rows]), want_t)
    with pytest.raises(ValueError):
        model.train().encode_codes(tblocks, space_to_depth="transposed")
