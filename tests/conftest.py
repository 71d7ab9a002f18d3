import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built C-ABI library; anywhere else they are
    skipped (with the reason), so a CPU-only run of the whole suite stays readable."""
    import torch
    from interactive_spectrogram_inpainting_b200 import _lib
    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    elif not _lib.library_path().exists():
        reason = f"{_lib.library_path().name} not built (python -m interactive_spectrogram_inpainting_b200.build)"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
