"""The product's VQVAE wiring (torch convs around the quantiser) against the unmodified
reference VQVAE on the CPU, with the oracle quantiser plugged into ours."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE
from oracle import ref_loader
from oracle.quantizer_oracle import OracleBottleneck


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("factors", [{'bottom': 16, 'top': 2}, {'bottom': 8, 'top': 4},
                                     {'bottom': 4, 'top': 2}])
def test_state_dict_and_encode_match_reference(factors):
    RefVQVAE = ref_loader.load_reference_vqvae_class()
    torch.manual_seed(0)
    ref = RefVQVAE(in_channel=2, resolution_factors=factors,
                   adapt_quantized_durations=False).eval()
    ours = VQVAE(in_channel=2, resolution_factors=factors, adapt_quantized_durations=False,
                 bottleneck_cls=OracleBottleneck).eval()
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True), None
    x = torch.randn(2, 2, 256, 32)
    with torch.no_grad():
        r = ref.encode(x.clone())
        o = ours.encode(x.clone())
        assert torch.equal(r[3], o[3]) and torch.equal(r[4], o[4])
        torch.testing.assert_close(r[0], o[0])
        torch.testing.assert_close(r[1], o[1])
        torch.testing.assert_close(r[2], o[2])
        rd = ref.decode_code(r[3], r[4])
        od = ours.decode_code(o[3], o[4])
        torch.testing.assert_close(rd, od)
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())


def test_fused_stack_walker_is_the_same_function(monkeypatch):
    """The inference path that folds bias/ReLU/skip-add into the convolutions (cuDNN fused ops
    on CUDA) must compute exactly what the module stack computes: run its control flow on
    the CPU with the two fused calls replaced by their definitions."""
    import torch.nn.functional as F
    from interactive_spectrogram_inpainting_b200.vqvae import vqvae as mod

    def conv(c, x, tf=None):
        return F.conv2d(x, mod._w(c, tf), c.bias, mod._hw(c.stride, tf), mod._hw(c.padding, tf),
                        mod._hw(c.dilation, tf), c.groups)

    monkeypatch.setattr(mod, "_can_fuse", lambda x: True)
    monkeypatch.setattr(mod, "_conv_relu", lambda c, x, tf=None: torch.relu(conv(c, x, tf)))
    monkeypatch.setattr(mod, "_conv_add_relu", lambda c, x, skip, tf=None: torch.relu(conv(c, x, tf) + skip))
    torch.manual_seed(3)
    for factor, n_res in ((16, 2), (2, 2), (4, 0), (8, 1)):
        enc = mod.Encoder(2, 32, n_res, 8, factor).eval()
        dec = mod.Decoder(32, 2, 32, n_res, 8, factor).eval()
        x = torch.randn(2, 2, 64, 32)
        with torch.no_grad():
            want = enc.blocks(x)
            got = enc(x)
            torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
            torch.testing.assert_close(dec(got), dec.blocks(want), rtol=1e-6, atol=1e-6)
            # the same walk on the transposed plane
            tf = mod.TransposedFilters()
            got_t = enc(x.transpose(2, 3).contiguous(), transposed=tf)
            torch.testing.assert_close(got_t.transpose(2, 3), want, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(dec(got_t, transposed=tf).transpose(2, 3), dec.blocks(want), rtol=1e-5, atol=1e-6)
    # a ResBlock that does not follow a convolution still rectifies its own input
    stack = torch.nn.Sequential(mod.ResBlock(4, 2), torch.nn.ReLU()).eval()
    x = torch.randn(1, 4, 8, 8)
    with torch.no_grad():
        torch.testing.assert_close(mod._run_blocks(stack, x), stack(x), rtol=1e-6, atol=1e-6)


def test_space_to_depth_first_conv_is_the_same_convolution():
    """The 4x4 / stride-2 first convolution regrouped as 3x3 / stride 1 over 2x2 input blocks
    (the layout the front-end kernel can write directly) gives the same encoder output, codes
    and gradients."""
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import SpectrogramsHelper
    from interactive_spectrogram_inpainting_b200.vqvae import vqvae as mod
    torch.manual_seed(4)
    for factor in (16, 8, 4, 2):
        enc = mod.Encoder(2, 32, 1, 8, factor).eval()
        x = torch.randn(2, 2, 64, 32)
        blocks = SpectrogramsHelper.to_space_to_depth(x)
        assert torch.equal(SpectrogramsHelper.from_space_to_depth(blocks), x)
        with torch.no_grad():
            torch.testing.assert_close(enc(blocks, space_to_depth=True), enc(x), rtol=1e-5, atol=1e-6)
    enc = mod.Encoder(2, 16, 0, 4, 4)
    x = torch.randn(1, 2, 16, 8)
    enc(x).square().sum().backward()
    want = enc.blocks[0].weight.grad.clone()
    enc.zero_grad()
    enc(SpectrogramsHelper.to_space_to_depth(x), space_to_depth=True).square().sum().backward()
    torch.testing.assert_close(enc.blocks[0].weight.grad, want, rtol=1e-4, atol=1e-5)
    model = VQVAE(in_channel=2, resolution_factors={'bottom': 16, 'top': 2},
                  bottleneck_cls=OracleBottleneck).eval()
    x = torch.randn(1, 2, 256, 32)
    with torch.no_grad():
        a = model.encode(x)
        b = model.encode(SpectrogramsHelper.to_space_to_depth(x), space_to_depth=True)
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    with pytest.raises(ValueError):
        mod.Encoder(2, 16, 0, 4, 4, use_local_kernels=True)(x, space_to_depth=True)


def test_every_reference_constructor_keyword_is_named_or_refused():
    """vqvae.py:66-98: nothing the reference dumps into model_parameters.json is dropped
    silently -- it is honoured, or the constructor raises."""
    import inspect
    ours = set(inspect.signature(VQVAE.__init__).parameters)
    reference_kwargs = {"encoders", "decoders", "in_channel", "num_hidden_channels", "n_res_block",
                        "num_residual_channels", "embed_dim", "num_embeddings", "decay", "groups",
                        "use_local_kernels", "output_activation_type", "output_spectrogram_min_magnitude",
                        "resolution_factors", "embeddings_initial_variance", "decoder_output_activation",
                        "normalizer_statistics", "corruption_weights", "adapt_quantized_durations",
                        "disable_quantization", "restarts_usage_threshold"}
    assert reference_kwargs <= ours
    assert not any(p.kind is inspect.Parameter.VAR_KEYWORD
                   for p in inspect.signature(VQVAE.__init__).parameters.values())
    kw = dict(in_channel=2, resolution_factors={'bottom': 16, 'top': 2}, bottleneck_cls=OracleBottleneck)
    for bad in (dict(restarts_usage_threshold=0.5), dict(groups=2), dict(encoders={'top': None, 'bottom': None}),
                dict(decoder_output_activation=torch.nn.ReLU()), dict(normalizer_statistics={'mean': 0., 'std': 1.})):
        with pytest.raises(NotImplementedError):
            VQVAE(**kw, **bad)
    with pytest.raises(AssertionError):
        VQVAE(**kw, output_activation_type='relu')
    with pytest.raises(TypeError):
        VQVAE(**kw, no_such_parameter=1)


def test_normaliser_and_masked_phase_output_transform(tmp_path):
    """vqvae.py:254-255 (normalise in ``encode``) and :297-302 (denormalise + masked-phase
    transform in ``decode``), with GANSynth's affine statistics; parameters survive the
    model_parameters.json round trip (vqvae.py:304-342)."""
    stats = {'s_a': 0.09, 's_b': 0.4, 'p_a': 0.8, 'p_b': 0.05}
    kw = dict(in_channel=2, resolution_factors={'bottom': 16, 'top': 2}, adapt_quantized_durations=False,
              bottleneck_cls=OracleBottleneck)
    torch.manual_seed(1)
    plain = VQVAE(**kw).eval()
    model = VQVAE(**kw, normalizer_statistics=stats, output_spectrogram_min_magnitude=-3.0).eval()
    model.load_state_dict(plain.state_dict())
    x = torch.randn(1, 2, 256, 32) * 3 - 4
    scale = torch.tensor([stats['s_a'], stats['p_a']]).view(1, 2, 1, 1)
    bias = torch.tensor([stats['s_b'], stats['p_b']]).view(1, 2, 1, 1)
    with torch.no_grad():
        got = model.encode(x)
        want = plain.encode(x * scale + bias)
        assert torch.equal(got[3], want[3]) and torch.equal(got[4], want[4])
        assert torch.equal(model.encode_codes(x)[1], want[4])
        raw = plain.decode(want[0], want[1])
        dec = model.decode(got[0], got[1])
        expect = (raw - bias) / scale
        expect[:, 1][expect[:, 0] < -3.0] = 0
        torch.testing.assert_close(dec, expect)
        assert (dec[:, 1][dec[:, 0] < -3.0] == 0).all() and (dec[:, 0] < -3.0).any()
        # the same affine as front-end knobs: helper applies it, the model must not re-apply it
        knobs = model.front_end_knobs()
        assert knobs["output_affine"] == ((0.09, 0.4), (0.8, 0.05))
        model.normalizes_input = False
        assert torch.equal(model.encode(x * scale + bias)[4], want[4])
        model.normalizes_input = True
        inv = model.inverse_front_end_knobs()["input_affine"]
        torch.testing.assert_close(raw[:, 0] * inv[0][0] + inv[0][1], expect[:, 0])
    path = tmp_path / "model_parameters.json"
    model.store_instantiation_parameters(path)
    import json
    params = json.loads(path.read_text())
    assert params["normalizer_statistics"] == stats and params["output_spectrogram_min_magnitude"] == -3.0
    again = VQVAE(**params, bottleneck_cls=OracleBottleneck)
    assert again.normalizer_statistics == stats and again.use_gansynth_normalization


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_unquantized_bottleneck_matches_reference():
    """bottleneck.py:107-119 / vqvae.py:159-160: ``disable_quantization`` passes the features
    through; runs on the CPU (no kernel involved), so the whole model is compared."""
    RefVQVAE = ref_loader.load_reference_vqvae_class()
    torch.manual_seed(2)
    kw = dict(in_channel=2, resolution_factors={'bottom': 16, 'top': 2}, adapt_quantized_durations=False,
              disable_quantization=True)
    ref = RefVQVAE(**kw).eval()
    ours = VQVAE(**kw).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(1, 2, 256, 32)
    with torch.no_grad():
        r, o = ref.encode(x), ours.encode(x)
    torch.testing.assert_close(r[0], o[0])
    torch.testing.assert_close(r[1], o[1])
    assert r[3] is None and o[3] is None and o[4] is None
    assert torch.isinf(o[5]).all() and float(o[2].sum()) == 0.0
    with pytest.raises(NotImplementedError):
        ours.quantize_t.embed_code(torch.zeros(1, 2, 2, dtype=torch.long))


def test_conv_stacks_on_the_transposed_plane_equal_the_plain_ones():
    """``TransposedFilters``: a conv stack run on ``x.transpose(2, 3)`` with every filter (and
    stride / padding pair) transposed gives the transposed result -- encoder (plain and
    space-to-depth first conv), decoder (with and without its last bias), CPU FP32."""
    import torch
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper as H
    from interactive_spectrogram_inpainting_b200.vqvae import vqvae as vq
    torch.manual_seed(3)
    tf = vq.TransposedFilters()
    x = torch.randn(2, 2, 64, 32)
    for factor in (16, 4, 2):
        enc = vq.Encoder(2, 64, 2, 16, factor).eval()
        with torch.no_grad():
            want = enc(x)
            got = enc(x.transpose(2, 3).contiguous(), transposed=tf).transpose(2, 3)
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
            blocks = H.to_space_to_depth(x).transpose(2, 3).contiguous()
            got = enc(blocks, space_to_depth=True, transposed=tf).transpose(2, 3)
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    for factor in (16, 2):
        dec = vq.Decoder(32, 2, 64, 2, 16, factor).eval()
        q = torch.randn(2, 32, 6, 3)
        with torch.no_grad():
            for no_bias in (False, True):
                want = dec(q, without_last_bias=no_bias)
                got = dec(q.transpose(2, 3).contiguous(), without_last_bias=no_bias, transposed=tf).transpose(2, 3)
                torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    # one copy per weight version
    w = enc.blocks[0].weight
    first = tf.weight(w)
    assert tf.weight(w) is first
    with torch.no_grad():
        w.add_(1.0)
    assert tf.weight(w) is not first and torch.equal(tf.weight(w), w.detach().transpose(2, 3))
    # the round trip of the block layout, both planes
    spec = torch.randn(2, 2, 8, 6)
    assert torch.equal(H.from_space_to_depth(H.to_space_to_depth(spec)), spec)
    assert torch.equal(H.from_space_to_depth(H.to_space_to_depth(spec).transpose(2, 3), transposed=True), spec)
