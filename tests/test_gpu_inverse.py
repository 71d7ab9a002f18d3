"""GPU parity of the inverse front end (isi_melif_inverse, ``to_audio``) against the
(unpinned) CPU restatement, through the C ABI."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import (
    MelSpectrogramsHelper, SpectrogramsHelper)
from oracle import frontend_oracle as fo
from test_imelif_emulation import _random_spec, check_audio_against_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _helper(mel, **kw):
    return (MelSpectrogramsHelper if mel else SpectrogramsHelper)(**kw).to(DEV)


@pytest.mark.parametrize("mel", [True, False])
def test_nsynth_shape_matches_oracle(mel):
    cfg = fo.FrontEndConfig(use_mel_scale=mel)
    audio = synthetic.synthetic_notes(3)
    spec = fo.to_spectrogram(audio.double(), cfg).float()
    got = _helper(mel).to_audio(spec.to(DEV))
    assert got.shape == (3, 64000) and got.dtype == torch.float32
    err = check_audio_against_oracle(got.cpu(), spec, cfg)
    print(f"[inverse] mel={mel}: max error {err:.2e} of the signal's max-abs")


@pytest.mark.parametrize("mel", [True, False])
@pytest.mark.parametrize("seg_frames", [8, 16, 48])
def test_segments_reproduce_the_whole_note(mel, seg_frames):
    cfg = fo.FrontEndConfig(use_mel_scale=mel)
    spec = _random_spec(2, 1024, 128, seed=3)
    helper = _helper(mel)
    helper.inverse_seg_frames = 128
    whole = helper.to_audio(spec.to(DEV))
    helper.inverse_seg_frames = seg_frames
    got = helper.to_audio(spec.to(DEV))
    check_audio_against_oracle(got.cpu(), spec, cfg)
    assert (got - whole).abs().max() <= 2e-5 * whole.abs().max()


@pytest.mark.parametrize("n_fft,hop,frames", [(1024, 256, 36), (512, 128, 30), (2048, 510, 21),
                                              (1024, 256, 17), (512, 512, 9), (2048, 512, 6)])
@pytest.mark.parametrize("mel", [True, False])
def test_other_geometries(n_fft, hop, frames, mel):
    """Unaligned hops (scalar overlap-add), frame counts that are not a multiple of four
    (synchronous slab fill, partial last batch), hop = n_fft (no overlap)."""
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft, use_mel_scale=mel)
    helper = _helper(mel, n_fft=n_fft, hop_length=hop, window_length=n_fft)
    spec = _random_spec(2, n_fft // 2, frames, seed=n_fft + frames)
    for seg in (None, 8):
        helper.inverse_seg_frames = seg
        check_audio_against_oracle(helper.to_audio(spec.to(DEV)).cpu(), spec, cfg)


def test_knobs_nyquist_bin_padding_and_affine():
    cfg = fo.FrontEndConfig(use_mel_scale=True, drop_bin="nyquist", pad_left=1024)
    helper = _helper(True, drop_bin="nyquist", pad_left=1024)
    helper.input_affine = ((2.0, -1.0), (0.5, 0.1))
    spec = _random_spec(1, 1024, 24, seed=11)
    check_audio_against_oracle(helper.to_audio(spec.to(DEV)).cpu(), spec, cfg, input_affine=helper.input_affine)


def test_batch_sizes_and_independence():
    """flask_server batch sizes 1-16 and a 444-note batch (the extraction step's size) choose
    different segmentations.  Every note equals the same note alone up to the segmentation's
    rounding, and within one launch equal inputs give bit-identical outputs."""
    helper = _helper(True)
    spec = _random_spec(2, 1024, 128, seed=21).to(DEV)
    alone = helper.to_audio(spec[:1])
    for b in (2, 5, 16):
        out = helper.to_audio(spec[:1].expand(b, -1, -1, -1))
        assert (out - alone).abs().max() <= 2e-5 * alone.abs().max()
        assert torch.equal(out[0], out[-1])
    big = helper.to_audio(spec.repeat(222, 1, 1, 1))
    assert big.shape == (444, 64000) and torch.isfinite(big).all()
    assert torch.equal(big[0], big[442]) and torch.equal(big[1], big[443])
    helper.inverse_seg_frames = 128
    whole = helper.to_audio(spec)
    assert (whole - big[:2]).abs().max() <= 2e-5 * whole.abs().max()


def test_round_trip_through_both_kernels():
    """audio -> isi_melif_forward -> isi_melif_inverse (linear scale): the audio comes back up to
    the eps of log(|X| + eps) and the dropped DC bin."""
    helper = _helper(False)
    audio = synthetic.synthetic_notes(4).to(DEV)
    back = helper.to_audio(helper.to_spectrogram(audio))
    assert back.shape == audio.shape
    assert (back - audio).abs().max() < 5e-3


def test_strided_and_3d_inputs_empty_batches_and_errors():
    helper = _helper(True)
    spec = _random_spec(2, 1024, 16, seed=2).to(DEV)
    ref = helper.to_audio(spec)
    strided = torch.empty(2, 2, 1024, 32, device=DEV)[..., ::2]
    strided.copy_(spec)
    assert torch.equal(helper.to_audio(strided), ref)
    assert torch.equal(helper.to_audio(spec[0])[0], helper.to_audio(spec[:1])[0])
    assert helper.to_audio(spec[:0]).shape == (0, 16 * 512 - 1536)
    with pytest.raises(RuntimeError, match="no CPU"):
        helper.to_audio(spec.cpu())
    with pytest.raises(ValueError):
        helper.to_audio(spec[:, :, :512])
    with pytest.raises(RuntimeError, match="differentiable"):
        helper.to_audio(spec.clone().requires_grad_(True))
    got = helper.to_audio_differentiable(spec)
    assert (got - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_decode_graph_continues_into_the_inverse_front_end():
    """The server's request: edited code maps -> decode_code -> to_audio, replayed from one
    CUDA graph (flask_server.py:593-596), equals the eager calls."""
    from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE, GraphedDecodeCode
    torch.manual_seed(0)
    model = VQVAE(in_channel=2, resolution_factors={'bottom': 16, 'top': 2},
                  adapt_quantized_durations=False).to(DEV).eval()
    helper = _helper(True)
    top, bottom = synthetic.synthetic_codemaps(2)
    top, bottom = top.to(DEV), bottom.to(DEV)
    graphed = GraphedDecodeCode(model, top, bottom, to_audio=helper)
    for shift in (0, 7):
        t, b = (top + shift) % 512, (bottom + shift) % 512
        spec, audio = graphed(t, b)
        with torch.no_grad():
            want_spec = model.decode_code(t, b)
        assert torch.allclose(spec, want_spec, rtol=1e-3, atol=1e-4)
        assert audio.shape == (2, 64000) and torch.isfinite(audio).all()
        assert torch.equal(audio, helper.to_audio(spec))      # same launch shape: bit-identical


def test_committed_golden_vectors(golden_dir):
    """tests/golden/inverse_unpinned.npz (oracle/make_golden.py::inverse_fixture)."""
    import numpy as np
    g = np.load(golden_dir / "inverse_unpinned.npz")
    gen = torch.Generator().manual_seed(20200117)
    spec = torch.stack([torch.randn(2, 1024, 24, generator=gen) * 2.0 - 3.0,
                        torch.rand(2, 1024, 24, generator=gen) * 2.0 - 1.0], 1)
    for mel, key in ((True, "mel"), (False, "lin")):
        got = _helper(mel).to_audio(spec.to(DEV))[:, ::7].cpu().numpy()
        assert np.abs(got - g[key]).max() <= 1e-4 * np.abs(g[key]).max()
