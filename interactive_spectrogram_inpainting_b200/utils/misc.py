"""Factory with the reference's signature (interactive_spectrogram_inpainting/utils/misc.py:10-29)."""
from .spectrograms_helper import MelSpectrogramsHelper, SpectrogramsHelper

# training-parameter key (as train_vqvae.py dumps it to command_line_parameters.json,
# utils/misc.py:13-27) -> constructor keyword of the helpers
_COMMON = ("fs_hz", "n_fft", "hop_length", "window_length")
_MEL = {"mel_scale_lower_edge_hertz": "lower_edge_hertz",
        "mel_scale_upper_edge_hertz": "upper_edge_hertz",
        "mel_scale_break_frequency_hertz": "mel_break_frequency_hertz",
        "mel_scale_expand_resolution_factor": "mel_bin_width_threshold_factor"}


def get_spectrograms_helper(**kwargs) -> SpectrogramsHelper:
    """Build the helper from a training-parameters dict; the mel variant iff ``use_mel_scale``.
    Keys the helpers do not take (the dict holds every command-line parameter) are ignored;
    a missing required key raises ``KeyError`` like the reference."""
    args = {k: kwargs[k] for k in _COMMON}
    if not kwargs["use_mel_scale"]:
        return SpectrogramsHelper(**args)
    args.update({ctor: kwargs[key] for key, ctor in _MEL.items()})
    return MelSpectrogramsHelper(**args)


def expand_path(p):
    """``utils/misc.py:32-33`` of the reference: user-expanded absolute path (the reference's
    scripts import it from this module)."""
    import pathlib
    return pathlib.Path(p).expanduser().absolute()
