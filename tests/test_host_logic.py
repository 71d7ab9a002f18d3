"""Host-side logic that needs no GPU: shard arithmetic, layout detection, band tables."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from interactive_spectrogram_inpainting_b200 import _lib
from interactive_spectrogram_inpainting_b200.utils import distributed as du
from interactive_spectrogram_inpainting_b200.utils import spectrograms_helper as sh


@settings(max_examples=200, deadline=None)
@given(total=st.integers(0, 5000), world=st.integers(1, 16))
def test_shards_partition_the_notes_exactly(total, world):
    seen_strided, seen_contig = [], []
    for rank in range(world):
        idx = list(du.shard_indices(total, rank, world))
        lo, hi = du.shard_range(total, rank, world)
        assert len(idx) == hi - lo == du.shard_size(total, rank, world)
        seen_strided += idx
        seen_contig += list(range(lo, hi))
    assert sorted(seen_strided) == list(range(total))      # no pad, no drop, no duplicate
    assert seen_contig == list(range(total))


@settings(max_examples=100, deadline=None)
@given(b=st.integers(1, 4), h=st.integers(1, 6), w=st.integers(1, 6), d=st.integers(1, 9))
def test_rows_layout_addresses_every_element(b, h, w, d):
    base = torch.arange(b * d * h * w, dtype=torch.float32).view(b, d, h, w)
    for view in (base.permute(0, 2, 3, 1), base.permute(0, 2, 3, 1).contiguous()):
        lay = _lib.rows_layout(view)
        assert lay is not None
        flat = view.reshape(-1, d)
        storage = view.contiguous().view(-1) if view.is_contiguous() else base.view(-1)
        for row in range(flat.shape[0]):
            bi, r = divmod(row, lay.rows_per_batch)
            for col in range(d):
                off = bi * lay.batch_stride + r * lay.row_stride + col * lay.col_stride
                assert storage[off] == flat[row, col]


def test_rows_layout_rejects_what_it_cannot_describe():
    t = torch.zeros(4, 6, 8, 16)[:, ::2, ::3]
    lay = _lib.rows_layout(t)
    if lay is not None:          # if described, it must be exact
        flat = t.reshape(-1, 16)
        assert lay.rows_per_batch * (flat.shape[0] // lay.rows_per_batch) == flat.shape[0]


@pytest.mark.parametrize("n_fft", [512, 1024, 2048])
def test_band_table_rows_are_contiguous_and_narrow(n_fft):
    starts, counts, weights = sh.mel_band_table(n_fft, 16000, 0.0, 8000.0, 700.0, 1.5)
    assert weights.shape[1] <= 8 and (counts <= weights.shape[1]).all()
    assert (starts + counts <= n_fft // 2).all() and (weights >= 0).all()
    for j in range(n_fft // 2):
        assert (weights[j, counts[j]:] == 0).all()


def test_helper_metadata_matches_reference_call_sites():
    h = sh.MelSpectrogramsHelper()
    assert (h.fs_hz, h.n_fft, h.hop_length, h.window_length) == (16000, 2048, 512, 2048)
    assert h.safelog_eps == 1e-6 and h.num_frames(64000) == 128 and h.n_freq == 1024
    with pytest.raises(ValueError):
        sh.SpectrogramsHelper(n_fft=1000)


def test_differentiable_to_audio_inverts_the_linear_front_end_on_cpu():
    """``to_audio_differentiable`` is plain torch (the spectral losses need its gradient): check
    it against the oracle's forward transform.  ``to_audio`` itself is the CUDA kernel."""
    from interactive_spectrogram_inpainting_b200.utils import synthetic
    from oracle import frontend_oracle as fo
    audio = synthetic.synthetic_notes(2)
    spec = fo.to_spectrogram(audio.double(), fo.FrontEndConfig(use_mel_scale=False)).float()
    rebuilt = sh.SpectrogramsHelper().to_audio_differentiable(spec)
    assert rebuilt.shape == audio.shape
    err = (rebuilt - audio)[:, 2048:-2048].abs().max()
    assert err < 2e-3, err          # exp(log(|X| + eps)) keeps the 1e-6 offset; edges excluded


def test_to_audio_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU"):
        sh.SpectrogramsHelper().to_audio(torch.zeros(1, 2, 1024, 8))


def test_differentiable_to_audio_matches_the_oracle_and_has_a_gradient():
    from oracle import frontend_oracle as fo
    g = torch.Generator().manual_seed(5)
    spec = torch.stack([torch.randn(1, 1024, 12, generator=g) - 3, torch.rand(1, 1024, 12, generator=g) * 2 - 1], 1)
    for mel in (True, False):
        helper = (sh.MelSpectrogramsHelper if mel else sh.SpectrogramsHelper)()
        want = fo.to_audio(spec.double(), fo.FrontEndConfig(use_mel_scale=mel))
        x = spec.clone().requires_grad_(True)
        got = helper.to_audio_differentiable(x)
        assert (got.double() - want).abs().max() <= 1e-5 * want.abs().max()
        got.pow(2).sum().backward()
        assert torch.isfinite(x.grad).all() and x.grad.abs().max() > 0


def test_mel_to_audio_runs_and_is_close_in_level():
    from interactive_spectrogram_inpainting_b200.utils import synthetic
    from oracle import frontend_oracle as fo
    audio = synthetic.synthetic_notes(1)
    spec = fo.to_spectrogram(audio, fo.FrontEndConfig())
    rebuilt = sh.MelSpectrogramsHelper().to_audio_differentiable(spec)
    assert rebuilt.shape == audio.shape and torch.isfinite(rebuilt).all()
    ratio = rebuilt.pow(2).mean().sqrt() / audio.pow(2).mean().sqrt()
    assert 0.3 < ratio < 3.0, ratio   # the mel pseudo-inverse is lossy; level must survive


def test_row_records_unpickle_as_the_reference_code_row():
    """extract.row_record writes the bytes extract_code.py:71-79 stores: key = note name, value =
    a pickle naming the reference's own CodeRow class, so the reference's readers
    (lmdb_dataset.py:79-89) load it without this package; write_lmdb commits a batch in one
    transaction."""
    import pickle
    import sys
    import types
    import numpy as np
    from interactive_spectrogram_inpainting_b200 import extract
    row = extract.CodeRow(top=np.arange(128).reshape(32, 4), bottom=np.arange(512).reshape(64, 8),
                          attributes={"pitch": 60}, filename="bass_synthetic_000-060-100")
    key, value = extract.row_record(row)
    assert key == b"bass_synthetic_000-060-100"
    ref_path = "interactive_spectrogram_inpainting.utils.datasets.lmdb_dataset"
    assert ref_path.encode() in value and b"interactive_spectrogram_inpainting_b200" not in value
    assert ref_path not in sys.modules                     # the placeholder is gone again
    # a reader that only has the reference's class (here: a stand-in module at its import path)
    from collections import namedtuple
    stand_in = types.ModuleType(ref_path)
    stand_in.CodeRow = namedtuple('CodeRow', ['top', 'bottom', 'attributes', 'filename'])
    stand_in.CodeRow.__module__ = ref_path
    parents = []
    parts = ref_path.split(".")
    for i in range(1, len(parts)):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            parents.append(name)
    sys.modules[ref_path] = stand_in
    try:
        loaded = pickle.loads(value)
        assert type(loaded) is stand_in.CodeRow
        assert (loaded.top == row.top).all() and (loaded.bottom == row.bottom).all()
        assert loaded.attributes == {"pitch": 60} and loaded.filename == row.filename
        assert loaded.top.dtype == np.int64
        # with the reference module loaded, its own class is pickled
        key2, value2 = extract.row_record(row)
        assert type(pickle.loads(value2)) is stand_in.CodeRow
    finally:
        del sys.modules[ref_path]
        for name in parents:
            del sys.modules[name]

    class Txn:
        def __init__(self, store): self.store = store
        def __enter__(self): self.store["begins"] += 1; return self
        def __exit__(self, *exc): return False
        def put(self, k, v): self.store[k] = v

    class Env:
        def __init__(self): self.store = {"begins": 0}
        def begin(self, db=None, write=False): assert write; return Txn(self.store)

    env = Env()
    rows = [row, row._replace(filename="other")]
    assert extract.write_lmdb(rows, env) == 2
    assert env.store["begins"] == 1 and set(env.store) == {"begins", key, b"other"}


def test_helper_factory_takes_the_reference_parameter_dict():
    """utils/misc.py:10-29: the dict train_vqvae.py dumps (every command-line parameter) selects
    and configures the helper; unrelated keys are ignored, a missing one raises KeyError."""
    from interactive_spectrogram_inpainting_b200.utils.misc import get_spectrograms_helper
    params = dict(fs_hz=16000, n_fft=1024, hop_length=256, window_length=1024, use_mel_scale=True,
                  mel_scale_lower_edge_hertz=20.0, mel_scale_upper_edge_hertz=7000.0,
                  mel_scale_break_frequency_hertz=650.0, mel_scale_expand_resolution_factor=2.0,
                  batch_size=64, lr=3e-4)
    mel = get_spectrograms_helper(**params)
    assert type(mel) is sh.MelSpectrogramsHelper
    assert (mel.fs_hz, mel.n_fft, mel.hop_length, mel.window_length) == (16000, 1024, 256, 1024)
    assert (mel.lower_edge_hertz, mel.upper_edge_hertz, mel.mel_break_frequency_hertz,
            mel.mel_bin_width_threshold_factor) == (20.0, 7000.0, 650.0, 2.0)
    lin = get_spectrograms_helper(**dict(params, use_mel_scale=False))
    assert type(lin) is sh.SpectrogramsHelper and lin.n_fft == 1024
    with pytest.raises(KeyError):
        get_spectrograms_helper(fs_hz=16000, n_fft=1024, use_mel_scale=False)
