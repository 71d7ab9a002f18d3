"""The drop-in boundary exercised the way a maintainer of the reference would use it.

(i)  INTEGRATION.md section 2's two-line swap, applied to the UNMODIFIED reference ``VQVAE``
     (vqvae.py:16-17,152-181; imported from /root/reference or the staged baseline/_ref copy):
     the reference model, on cuda, with this repo's ``QuantizedBottleneck`` inside, against the
     same reference model on the CPU with its own bottleneck -- ``encode``, ``decode_code`` and
     a state-dict round trip.
(ii) The threaded-server contract (flask_server.py:296-299, SURVEY.md 8b "Threading"):
     ``forward`` / ``embed_code`` / ``to_spectrogram`` called from 4 threads on 4 CUDA streams
     concurrently return what the serial calls return, bit for bit."""
import threading

import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
from oracle import parity, ref_loader

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
MODEL_KW = dict(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)


@pytest.fixture
def fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.fixture
def patched_reference_vqvae():
    """INTEGRATION.md section 2, verbatim."""
    if not ref_loader.available():
        pytest.skip("neither /root/reference nor baseline/_ref is present")
    ref_cls = ref_loader.load_reference_vqvae_class()
    import interactive_spectrogram_inpainting.vqvae.vqvae as ref_vqvae
    stock = ref_vqvae.QuantizedBottleneck
    ref_vqvae.QuantizedBottleneck = QuantizedBottleneck
    try:
        yield ref_cls, stock
    finally:
        ref_vqvae.QuantizedBottleneck = stock


def test_reference_vqvae_with_the_b200_quantiser_inside(patched_reference_vqvae, fp32_convs, capsys):
    ref_cls, stock_bottleneck = patched_reference_vqvae
    import interactive_spectrogram_inpainting.vqvae.vqvae as ref_vqvae
    torch.manual_seed(21)
    gpu_model = ref_cls(**MODEL_KW)
    assert type(gpu_model.quantize_t) is QuantizedBottleneck and type(gpu_model.quantize_b) is QuantizedBottleneck
    # the CPU twin is the reference with its own bottleneck
    ref_vqvae.QuantizedBottleneck = stock_bottleneck
    cpu_model = ref_cls(**MODEL_KW).eval()
    ref_vqvae.QuantizedBottleneck = QuantizedBottleneck
    assert type(cpu_model.quantize_t) is stock_bottleneck
    # state-dict round trip, both directions, strict: same keys, same shapes (bottleneck.py:49-51)
    cpu_model.load_state_dict(gpu_model.state_dict(), strict=True)
    gpu_model.load_state_dict(cpu_model.state_dict(), strict=True)
    assert {k for k in gpu_model.state_dict() if k.startswith("quantize_t.")} == {
        "quantize_t.embed", "quantize_t.cluster_size", "quantize_t.embed_avg"}
    gpu_model = gpu_model.to(DEV).eval()

    from oracle import frontend_oracle as fo
    audio = synthetic.synthetic_notes(8)
    spec = fo.to_spectrogram(audio.double(), fo.FrontEndConfig()).float()
    with torch.no_grad():
        got = gpu_model.encode(spec.to(DEV))                 # vqvae.py:251-278 on cuda, our kernels inside
        want = cpu_model.encode(spec)
        feat_t, want_t, feat_b, want_b = parity.encode_with_features(cpu_model, spec)
        gfeat_t, _, gfeat_b, _ = parity.encode_with_features(gpu_model, spec.to(DEV))
    names = ("quant_t", "quant_b", "diff", "id_t", "id_b", "perplexity_t", "perplexity_b")
    for name, g, w in zip(names, got, want):
        assert g.shape == w.shape and g.dtype == w.dtype, name
    assert got[2].shape == (1,) and got[3].dtype == torch.int64
    assert torch.equal(want[3], want_t) and torch.equal(want[4], want_b)
    rep_t = parity.explain_differences(feat_t, got[3], want_t, cpu_model.quantize_t.embed, gfeat_t.cpu() - feat_t)
    same_top = (got[3].cpu() == want_t).reshape(8, -1).all(1)
    rep_b = parity.explain_differences(feat_b[same_top], got[4].cpu()[same_top], want_b[same_top],
                                       cpu_model.quantize_b.embed, (gfeat_b.cpu() - feat_b)[same_top])
    with capsys.disabled():
        print(f"\n[drop-in, reference VQVAE.encode on cuda] top: {rep_t}\n[drop-in] bottom: {rep_b}")
    assert rep_t.unexplained == 0 and rep_b.unexplained == 0
    if bool(same_top.all()) and torch.equal(got[4].cpu(), want_b):
        # identical code maps: the dequantised tensors and the scalars agree to FP32 rounding
        torch.testing.assert_close(got[0].cpu(), want[0], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(got[1].cpu(), want[1], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(got[2].cpu(), want[2], rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(got[5].cpu(), want[5], rtol=1e-4, atol=1e-4)

    # decode_code (vqvae.py:288-295): same code maps in, same spectrogram out
    top, bottom = synthetic.synthetic_codemaps(3)
    with torch.no_grad():
        dec_gpu = gpu_model.decode_code(top.to(DEV), bottom.to(DEV))
        dec_cpu = cpu_model.decode_code(top, bottom)
    assert dec_gpu.shape == dec_cpu.shape == (3, 2, 1024, 128)
    torch.testing.assert_close(dec_gpu.cpu(), dec_cpu, rtol=1e-3, atol=1e-4)
    # embed_code is a pure lookup: bit-exact
    assert torch.equal(gpu_model.quantize_b.embed_code(bottom.to(DEV)).cpu(), cpu_model.quantize_b.embed_code(bottom))

    # training mode keeps the module contract: buffers move in place, gradients reach the encoder
    gpu_model.train()
    before = gpu_model.quantize_t.embed.clone()
    out = gpu_model(spec[:2].to(DEV))                        # forward(): encode + decode, 6-tuple (vqvae.py:245-249)
    assert len(out) == 6 and out[0].shape == (2, 2, 1024, 128)
    (out[0].pow(2).mean() + out[1].sum()).backward()
    assert not torch.equal(before, gpu_model.quantize_t.embed)
    assert all(p.grad is not None for p in gpu_model.enc_b.parameters())


def test_four_threads_on_four_streams_equal_the_serial_calls():
    torch.manual_seed(3)
    n_threads = 4
    quantiser = QuantizedBottleneck(64, 512).to(DEV).eval()
    helper = MelSpectrogramsHelper().to(DEV)
    xs = [torch.randn(2, 32, 4, 64, device=DEV) * (0.5 + i) for i in range(n_threads)]
    ids = [torch.randint(0, 512, (1 + i, 64, 8), device=DEV) for i in range(n_threads)]
    audios = [synthetic.synthetic_notes(1 + i, seed=100 + i).to(DEV) for i in range(n_threads)]
    with torch.no_grad():
        serial = [(quantiser(x), quantiser.embed_code(i), helper.to_spectrogram(a))
                  for x, i, a in zip(xs, ids, audios)]
    torch.cuda.synchronize()
    results, errors = [None] * n_threads, []
    start = threading.Barrier(n_threads)

    def worker(k):
        try:
            stream = torch.cuda.Stream(DEV)
            stream.wait_stream(torch.cuda.default_stream(DEV))
            start.wait()
            with torch.no_grad(), torch.cuda.stream(stream):
                outs = []
                for _ in range(20):                          # many interleaved launches per thread
                    outs = (quantiser(xs[k]), quantiser.embed_code(ids[k]), helper.to_spectrogram(audios[k]))
                stream.synchronize()
            results[k] = outs
        except Exception as exc:                             # surfaced in the main thread
            errors.append((k, repr(exc)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(n_threads):
        (q, diff, ind, perp), emb, spec = results[k]
        (q0, diff0, ind0, perp0), emb0, spec0 = serial[k]
        assert torch.equal(ind, ind0) and torch.equal(q, q0) and torch.equal(emb, emb0)
        assert torch.equal(diff, diff0) and torch.equal(perp, perp0)
        assert torch.equal(spec, spec0)
