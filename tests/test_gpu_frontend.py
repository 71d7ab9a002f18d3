"""GPU parity of the fused front-end kernel against the (unpinned) CPU restatement."""
import pytest
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.utils.misc import get_spectrograms_helper
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import (
    MelSpectrogramsHelper, SpectrogramsHelper)
from oracle import frontend_oracle as fo
from test_melif_emulation import check_against_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("use_mel", [True, False])
def test_nsynth_shape_matches_oracle(use_mel):
    audio = synthetic.synthetic_notes(3)
    helper = get_spectrograms_helper(
        fs_hz=16000, n_fft=2048, hop_length=512, window_length=2048, use_mel_scale=use_mel,
        mel_scale_lower_edge_hertz=0.0, mel_scale_upper_edge_hertz=8000.0,
        mel_scale_break_frequency_hertz=700.0, mel_scale_expand_resolution_factor=1.5).to(DEV)
    spec = helper.to_spectrogram(audio.to(DEV))
    assert spec.shape == (3, 2, 1024, 128) and spec.dtype == torch.float32
    excluded = check_against_oracle(spec.cpu(), audio, fo.FrontEndConfig(use_mel_scale=use_mel))
    print(f"[front end] mel={use_mel}: {excluded:.3%} positions ill-conditioned in FP32")


@pytest.mark.parametrize("n_fft,hop,samples", [(1024, 256, 9000), (512, 128, 4099),
                                               (2048, 512, 300), (2048, 512, 96001),
                                               # other hops on the warp-specialised kernel (n_fft 2048,
                                               # sample count and hop multiples of 8)
                                               (2048, 256, 16000), (2048, 1024, 32000), (2048, 128, 8000)])
def test_other_sizes_and_ragged_lengths(n_fft, hop, samples):
    audio = synthetic.synthetic_notes(2, n_samples=samples)
    helper = MelSpectrogramsHelper(n_fft=n_fft, hop_length=hop, window_length=n_fft).to(DEV)
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    spec = helper.to_spectrogram(audio.to(DEV))
    assert spec.shape[-1] == fo.frame_geometry(cfg, samples)[2]
    check_against_oracle(spec.cpu(), audio, cfg)


def test_knobs_nyquist_symmetric_window():
    audio = synthetic.synthetic_notes(1, n_samples=16000)
    helper = SpectrogramsHelper(drop_bin="nyquist", window_periodic=False).to(DEV)
    cfg = fo.FrontEndConfig(use_mel_scale=False, drop_bin="nyquist", window_periodic=False)
    check_against_oracle(helper.to_spectrogram(audio.to(DEV)).cpu(), audio, cfg)


def test_silence_and_batch_independence():
    helper = MelSpectrogramsHelper().to(DEV)
    audio = synthetic.synthetic_notes(5).to(DEV)
    audio[2] = 0
    spec = helper.to_spectrogram(audio)
    assert torch.isfinite(spec).all()
    assert torch.allclose(spec[2, 0], torch.full_like(spec[2, 0], float(torch.log(torch.tensor(1e-6)))))
    assert (spec[2, 1] == 0).all()
    alone = helper.to_spectrogram(audio[3:4])
    assert torch.equal(alone[0], spec[3])          # notes do not interact; deterministic


def test_full_batch_size_property():
    """cfg 2 batch (256 notes): every note equals the same note run alone."""
    helper = MelSpectrogramsHelper().to(DEV)
    audio = synthetic.synthetic_notes(256).to(DEV)
    spec = helper.to_spectrogram(audio)
    for i in (0, 100, 255):
        assert torch.equal(helper.to_spectrogram(audio[i:i + 1])[0], spec[i])


def test_channels_last_storage_holds_the_same_tensor():
    audio = synthetic.synthetic_notes(3).to(DEV)
    plain = MelSpectrogramsHelper().to(DEV).to_spectrogram(audio)
    nhwc = MelSpectrogramsHelper(channels_last=True).to(DEV).to_spectrogram(audio)
    assert nhwc.shape == plain.shape and nhwc.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(nhwc, plain)
    lin = SpectrogramsHelper(channels_last=True).to(DEV).to_spectrogram(audio[:, :9001])
    assert torch.equal(lin, SpectrogramsHelper().to(DEV).to_spectrogram(audio[:, :9001]))


def test_small_batches_are_split_into_frame_segments_identically():
    """B=1 runs as several frame segments per note (look-back transform at each seam);
    B=300 runs whole notes: the two must agree bit for bit."""
    helper = MelSpectrogramsHelper().to(DEV)
    audio = synthetic.synthetic_notes(300).to(DEV)
    whole = helper.to_spectrogram(audio)
    for i in (0, 7, 299):
        assert torch.equal(helper.to_spectrogram(audio[i:i + 1])[0], whole[i])


def test_fused_masked_phase_and_normaliser_epilogue():
    audio = synthetic.synthetic_notes(2)
    kw = dict(masked_phase_threshold=-3.0, output_affine=((0.1, 0.5), (2.0, -0.25)))
    for cl in (False, True):
        plain = MelSpectrogramsHelper(channels_last=cl).to(DEV).to_spectrogram(audio.to(DEV))
        fused = MelSpectrogramsHelper(channels_last=cl, **kw).to(DEV).to_spectrogram(audio.to(DEV))
        want = fo.epilogue(plain.cpu(), **kw)
        assert torch.allclose(fused.cpu(), want, rtol=0, atol=1e-6)
        assert (fused[:, 1][plain[:, 0] < -3.0] == -0.25).all()


def test_band_table_pitch_and_dropped_nyquist_in_mel_mode():
    """The kernel reads 8-wide band rows with two 16-byte loads and any other pitch with scalar
    loads: both give the same bits.  Mel mode with the Nyquist bin dropped exercises the
    branch-free item 0 of the polar step with the DC bin as its real bin."""
    audio = synthetic.synthetic_notes(2, n_samples=20000)
    helper = MelSpectrogramsHelper().to(DEV)
    padded = helper.to_spectrogram(audio.to(DEV))
    width = int(helper.mel_count.max())
    assert helper.mel_weight.shape[1] == 8 and width < 8
    helper.mel_weight = helper.mel_weight[:, :width].contiguous()
    assert torch.equal(helper.to_spectrogram(audio.to(DEV)), padded)

    helper = MelSpectrogramsHelper(drop_bin="nyquist").to(DEV)
    cfg = fo.FrontEndConfig(drop_bin="nyquist")
    check_against_oracle(helper.to_spectrogram(audio.to(DEV)).cpu(), audio, cfg)


@pytest.mark.parametrize("n_fft,hop,samples", [(2048, 512, 64000), (2048, 512, 9999), (1024, 256, 8001),
                                               (512, 128, 4096)])
def test_pcm16_input_equals_float_input(n_fft, hop, samples):
    """int16 PCM audio (NSynth's storage format) is converted inside the kernel,
    float(x) * pcm_scale: the same bits as converting first and uploading FP32 -- on the
    bulk-copy path (aligned sizes) and on the per-thread staging path (ragged sizes)."""
    audio = synthetic.synthetic_notes(3, n_samples=samples)
    pcm = (audio * 32767.0).round().clamp(-32768, 32767).to(torch.int16)
    helper = MelSpectrogramsHelper(n_fft=n_fft, hop_length=hop, window_length=n_fft).to(DEV)
    as_float = pcm.to(DEV).float() * helper.pcm_scale
    want = helper.to_spectrogram(as_float)
    got = helper.to_spectrogram(pcm.to(DEV))
    assert got.dtype == torch.float32 and torch.equal(got, want)
    helper.pcm_scale = 1.0 / 32767.0
    assert torch.equal(helper.to_spectrogram(pcm.to(DEV)),
                       helper.to_spectrogram(pcm.to(DEV).float() * torch.tensor(1.0 / 32767.0, device=DEV)))
    # and the float result is the oracle's
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    check_against_oracle(want.cpu(), as_float.cpu(), cfg)


@pytest.mark.parametrize("samples", [64000, 9001])
def test_space_to_depth_output_is_the_same_spectrogram(samples):
    """layout 2: 2x2 blocks of the spectrogram as 8 channels, written directly by the kernel
    (whole notes, frame segments, the ragged last batch and the fused epilogue included)."""
    audio = synthetic.synthetic_notes(3, n_samples=samples).to(DEV)
    kw = dict(masked_phase_threshold=-10.0, output_affine=((0.1, 0.5), (2.0, -0.25)))
    for extra in ({}, kw):
        frames = MelSpectrogramsHelper().num_frames(samples)
        n_frames = frames + (frames % 2)                       # the layout needs an even count
        plain = MelSpectrogramsHelper(n_frames=n_frames, **extra).to(DEV).to_spectrogram(audio)
        blocks = MelSpectrogramsHelper(n_frames=n_frames, space_to_depth=True, **extra).to(DEV).to_spectrogram(audio)
        assert blocks.shape == (3, 8, 512, n_frames // 2)
        assert blocks.permute(0, 2, 3, 1).is_contiguous()
        assert torch.equal(MelSpectrogramsHelper.from_space_to_depth(blocks), plain)
        # layout 3: the same blocks on the transposed plane (frequency fastest in memory)
        tblocks = MelSpectrogramsHelper(n_frames=n_frames, space_to_depth="transposed", **extra).to(DEV).to_spectrogram(audio)
        assert tblocks.shape == (3, 8, n_frames // 2, 512)
        assert tblocks.permute(0, 2, 3, 1).is_contiguous()
        assert torch.equal(tblocks.transpose(2, 3), blocks)
        assert torch.equal(MelSpectrogramsHelper.from_space_to_depth(tblocks, transposed=True), plain)
    with pytest.raises(ValueError):
        MelSpectrogramsHelper(n_frames=127, space_to_depth=True).to(DEV).to_spectrogram(audio)
    with pytest.raises(ValueError):
        MelSpectrogramsHelper(space_to_depth="sideways")


@pytest.mark.parametrize("n_frames", [128, 126, 124, 121])
def test_every_layout_holds_the_same_values_at_any_frame_count(n_frames):
    """The planar and channels-last stores take 32-byte vectors when the row pitch allows
    (n_frames % 8 / % 4) and 16-byte or scalar stores otherwise; an unaligned view of the output
    buffer must not matter either (the helper allocates, so alignment is torch's: 256 bytes)."""
    audio = synthetic.synthetic_notes(5, n_samples=61000).to(DEV)            # fits in 121 frames
    plain = MelSpectrogramsHelper(n_frames=n_frames).to(DEV).to_spectrogram(audio)
    assert plain.shape == (5, 2, 1024, n_frames) and plain.is_contiguous()
    cl = MelSpectrogramsHelper(n_frames=n_frames, channels_last=True).to(DEV).to_spectrogram(audio)
    assert cl.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(cl, plain)
    reference = MelSpectrogramsHelper(n_frames=128).to(DEV).to_spectrogram(audio)
    assert torch.equal(plain[..., :121], reference[..., :121])
    if n_frames % 2 == 0:
        for mode in (True, "transposed"):
            blocks = MelSpectrogramsHelper(n_frames=n_frames, space_to_depth=mode).to(DEV).to_spectrogram(audio)
            assert torch.equal(MelSpectrogramsHelper.from_space_to_depth(blocks, transposed=mode == "transposed"), plain)


@pytest.mark.parametrize("n_fft,hop,samples", [(512, 125, 3001), (1024, 250, 5000), (2048, 500, 40000)])
def test_odd_and_unaligned_hops_take_the_scalar_paths(n_fft, hop, samples):
    """Hops that are odd or not a multiple of the 16-byte bulk-copy granule: per-thread staging
    and scalar sample loads, for FP32 and for PCM input."""
    audio = synthetic.synthetic_notes(2, n_samples=samples)
    helper = MelSpectrogramsHelper(n_fft=n_fft, hop_length=hop, window_length=n_fft).to(DEV)
    cfg = fo.FrontEndConfig(n_fft=n_fft, hop_length=hop, window_length=n_fft)
    spec = helper.to_spectrogram(audio.to(DEV))
    check_against_oracle(spec.cpu(), audio, cfg)
    pcm = (audio * 32767.0).round().to(torch.int16)
    assert torch.equal(helper.to_spectrogram(pcm.to(DEV)),
                       helper.to_spectrogram(pcm.to(DEV).float() * helper.pcm_scale))


def test_empty_batch_and_very_short_audio():
    helper = MelSpectrogramsHelper().to(DEV)
    empty = helper.to_spectrogram(torch.zeros(0, 64000, device=DEV))
    assert empty.shape == (0, 2, 1024, 128)
    short = synthetic.synthetic_notes(2, n_samples=300)
    spec = helper.to_spectrogram(short.to(DEV))
    assert spec.shape[:3] == (2, 2, 1024) and torch.isfinite(spec).all()
    check_against_oracle(spec.cpu(), short, fo.FrontEndConfig())


def test_repeated_launches_are_bit_identical_under_load():
    """The warp-specialised kernel hands workspaces and audio stages between its two roles
    through mbarriers (racecheck does not model them and reports the hand-offs as hazards):
    a lost or early hand-off would change bits.  40 launches of a full 444-note batch, and the
    same notes inside batches of other sizes (other segmentations), must reproduce the first
    result exactly."""
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    helper = MelSpectrogramsHelper(space_to_depth=True).to(DEV)
    base = synthetic.synthetic_notes(64)
    audio = torch.cat([torch.roll(base, 997 * r, 1) for r in range(7)])[:444]
    pcm = (audio * 32767).round().to(torch.int16).to(DEV)
    first = helper.to_spectrogram(pcm).clone()
    for _ in range(40):
        assert torch.equal(helper.to_spectrogram(pcm), first)
    for n in (1, 5, 37, 148, 300):
        assert torch.equal(helper.to_spectrogram(pcm[:n]), first[:n]), n
    plain = MelSpectrogramsHelper().to(DEV)
    assert torch.equal(MelSpectrogramsHelper.from_space_to_depth(first[:9]), plain.to_spectrogram(pcm[:9]))


@pytest.mark.parametrize("layout", ["planar", "channels_last", "blocks", "blocks_t"])
def test_output_buffer_aligned_to_16_but_not_32_bytes(layout):
    """C-ABI callers own the output buffer: 16-byte alignment is the contract (ISI_ERR_ALIGN below
    it); the 32-byte stores must step aside for a buffer that is only 16-byte aligned."""
    from interactive_spectrogram_inpainting_b200 import _lib
    kw = {"planar": {}, "channels_last": dict(channels_last=True), "blocks": dict(space_to_depth=True),
          "blocks_t": dict(space_to_depth="transposed")}[layout]
    helper = MelSpectrogramsHelper(**kw).to(DEV)
    pcm = (synthetic.synthetic_notes(3) * 32767).round().to(torch.int16).to(DEV)
    want = helper.to_spectrogram(pcm)
    params = helper._params(128)
    params.audio_format, params.pcm_scale = _lib.AUDIO_PCM16, helper.pcm_scale
    raw = torch.zeros(want.numel() + 8, dtype=torch.float32, device=DEV)
    for offset, ok in ((4, True), (1, False)):
        out = raw[offset:offset + want.numel()]
        assert out.data_ptr() % 32 == (16 if ok else 4)
        if ok:
            _lib.invoke("isi_melif_forward", pcm.data_ptr(), 3, pcm.shape[1], params, out.data_ptr(),
                        _lib.stream_ptr(pcm.device))
            stored = want.permute(0, 2, 3, 1) if layout != "planar" else want       # memory order
            assert torch.equal(out.view(stored.shape), stored.contiguous() if layout == "planar" else stored)
        else:
            with pytest.raises(RuntimeError):
                _lib.invoke("isi_melif_forward", pcm.data_ptr(), 3, pcm.shape[1], params, out.data_ptr(),
                            _lib.stream_ptr(pcm.device))
