"""Factory with the reference's signature (interactive_spectrogram_inpainting/utils/misc.py:10-29)."""
import pathlib
from typing import Union

from .spectrograms_helper import MelSpectrogramsHelper, SpectrogramsHelper


def get_spectrograms_helper(**kwargs) -> SpectrogramsHelper:
    """Build the helper from a training-parameters dict (keys as dumped by
    train_vqvae.py into command_line_parameters.json: utils/misc.py:13-27)."""
    common = dict(fs_hz=kwargs['fs_hz'], n_fft=kwargs['n_fft'],
                  hop_length=kwargs['hop_length'], window_length=kwargs['window_length'])
    if kwargs['use_mel_scale']:
        return MelSpectrogramsHelper(
            **common,
            lower_edge_hertz=kwargs['mel_scale_lower_edge_hertz'],
            upper_edge_hertz=kwargs['mel_scale_upper_edge_hertz'],
            mel_break_frequency_hertz=kwargs['mel_scale_break_frequency_hertz'],
            mel_bin_width_threshold_factor=kwargs['mel_scale_expand_resolution_factor'])
    return SpectrogramsHelper(**common)


def expand_path(p: Union[str, pathlib.Path]) -> pathlib.Path:
    return pathlib.Path(p).expanduser().absolute()
