"""One training step with the semantics of the reference's ``train()`` (train_vqvae.py:169-192:
``model.train()``, ``zero_grad``, forward, reconstruction criterion + ``latent_loss_weight *
latent_loss.mean()``, backward, optimiser step) -- this repo's ``VQVAE`` on the GPU against the
UNMODIFIED reference ``VQVAE`` on the CPU with the same weights: codes, loss terms, the gradient
of every parameter, the EMA buffers of both quantisers (bottleneck.py:79-92) and the parameters
after the Adam step.  Plus the AMP variant of the step (train_vqvae.py:174, ``use_amp``)."""
import pytest
import torch
import torch.nn.functional as F

from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE
from oracle import ref_loader

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
MODEL_KW = dict(in_channel=2, resolution_factors={"bottom": 16, "top": 2}, adapt_quantized_durations=False)
LATENT_LOSS_WEIGHT = 0.25                      # train_vqvae.py:143


def _step(model, optimizer, img, use_amp=False):
    """train_vqvae.py:169-192 without the logging."""
    model.train()
    model.zero_grad()
    with torch.autocast(device_type=img.device.type, enabled=use_amp):
        out, latent_loss, perplexity_t, perplexity_b, id_t, id_b = model(img)
        reconstruction_loss = F.mse_loss(out, img)
    latent_loss = latent_loss.mean()
    loss = reconstruction_loss + LATENT_LOSS_WEIGHT * latent_loss
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    optimizer.step()
    return dict(loss=loss.detach(), rec=reconstruction_loss.detach(), latent=latent_loss.detach(),
                perplexity_t=perplexity_t.detach(), perplexity_b=perplexity_b.detach(),
                id_t=id_t, id_b=id_b, grads=grads)


@pytest.fixture
def fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.skipif(not ref_loader.available(), reason="neither /root/reference nor baseline/_ref is present")
def test_training_step_matches_the_reference_step(fp32_convs):
    RefVQVAE = ref_loader.load_reference_vqvae_class()
    torch.manual_seed(4)
    ref = RefVQVAE(**MODEL_KW)
    ours = VQVAE(**MODEL_KW)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours = ours.to(DEV)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=3e-4)
    opt_ours = torch.optim.Adam(ours.parameters(), lr=3e-4)
    g = torch.Generator().manual_seed(5)
    for step in range(2):
        img = torch.randn(2, 2, 256, 32, generator=g)
        want = _step(ref, opt_ref, img)
        got = _step(ours, opt_ours, img.to(DEV))
        # a near-tie flip would fork the two trajectories: with 80 rows per step it does not happen
        assert torch.equal(got["id_t"].cpu(), want["id_t"]) and torch.equal(got["id_b"].cpu(), want["id_b"])
        for key in ("loss", "rec", "latent", "perplexity_t", "perplexity_b"):
            torch.testing.assert_close(got[key].cpu().reshape(-1), want[key].reshape(-1), rtol=2e-4, atol=1e-6)
        assert set(got["grads"]) == set(want["grads"])
        for name, grad in want["grads"].items():
            err = (got["grads"][name].cpu() - grad).abs().max()
            assert err <= 2e-3 * grad.abs().max() + 1e-7, (step, name, float(err), float(grad.abs().max()))
        state_ref, state_ours = ref.state_dict(), ours.state_dict()
        for name in state_ref:
            if name.startswith("quantize_t.") or name.startswith("quantize_b."):     # EMA buffers
                torch.testing.assert_close(state_ours[name].cpu(), state_ref[name], rtol=1e-4, atol=1e-6,
                                           msg=lambda m, n=name: f"{n}: {m}")
    # parameters after two Adam steps (Adam normalises the gradient: compare loosely, but compare)
    for name, p in ref.named_parameters():
        q = dict(ours.named_parameters())[name]
        assert (q.detach().cpu() - p.detach()).abs().max() <= 1e-3, name


def test_amp_training_step_runs_and_keeps_the_module_contract():
    """train_vqvae.py:174 with ``use_amp``: under autocast the pre-quantiser features arrive in
    FP16; the quantiser searches in FP32 on their values, returns ``quantize`` in the input's dtype,
    FP32 scalars and int64 codes, and gradients reach the encoder."""
    torch.manual_seed(6)
    model = VQVAE(**MODEL_KW).to(DEV)
    optimizer = torch.optim.Adam(model.parameters(), lr=3e-4)
    img = torch.randn(2, 2, 256, 32, device=DEV)
    out = _step(model, optimizer, img, use_amp=True)
    assert torch.isfinite(out["loss"]) and out["id_t"].dtype == torch.int64
    assert all(torch.isfinite(g).all() for g in out["grads"].values())
    assert any(k.startswith("enc_b.") for k in out["grads"])
    model.eval()
    feats = torch.randn(4, 8, 2, 64, device=DEV).half()
    with torch.no_grad(), torch.autocast(device_type="cuda"):
        quant, diff, ind, perp = model.quantize_t(feats)
    assert quant.dtype == torch.float16 and diff.dtype == torch.float32 and ind.dtype == torch.int64
    with torch.no_grad():
        _, _, ind32, _ = model.quantize_t(feats.float())       # the same values in FP32
    assert torch.equal(ind32, ind)
