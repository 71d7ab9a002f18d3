"""The C-ABI library loads and exports every symbol include/isi_b200.h declares.
No compute calls: this runs on the GPU-less build box."""
import ctypes
import pathlib
import re

import pytest

from interactive_spectrogram_inpainting_b200 import _lib, build

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = (ROOT / "include" / "isi_b200.h").read_text()
    declared = set(re.findall(r"ISI_API[^;(]*?\b(isi_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_status_strings_and_pure_host_queries(lib):
    assert lib.isi_version() >= 100
    assert lib.isi_status_string(0) == b"ok"
    assert b"NULL" in lib.isi_status_string(-1)
    assert lib.isi_vq_prepared_bytes(64, 512) >= 4 * (512 + 2 * 64 * 512)
    assert lib.isi_vq_prepared_bytes(0, 512) == 0
    assert lib.isi_vq_gather_workspace_bytes(1000, 64) % 8 == 0


def test_argument_validation_happens_before_any_launch(lib):
    lay = _lib.RowsLayout(4, 0, 64, 1)
    assert lib.isi_vq_assign(None, lay, 4, 64, 512, None, None, None, 0, None) == -1
    assert lib.isi_vq_prepare_codebook(None, 64, 512, None, 0, None) == -1
    assert lib.isi_melif_forward(None, 1, 64000, None, None, None) == -1


def test_cpu_tensors_are_refused_loudly():
    import torch
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import SpectrogramsHelper
    from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QuantizedBottleneck(64, 512)(torch.zeros(2, 4, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SpectrogramsHelper().to_spectrogram(torch.zeros(1, 64000))


def test_rows_layout_detection():
    import torch
    nchw = torch.zeros(3, 64, 5, 7)
    lay = _lib.rows_layout(nchw.permute(0, 2, 3, 1))
    assert (lay.rows_per_batch, lay.batch_stride, lay.row_stride, lay.col_stride) == (35, 2240, 1, 35)
    lay = _lib.rows_layout(torch.zeros(3, 5, 7, 64))
    assert (lay.rows_per_batch, lay.row_stride, lay.col_stride) == (105, 64, 1)
    # strides that are not 'batches of uniformly strided rows' are reported as such (-> copy)
    assert _lib.rows_layout(torch.zeros(4, 6, 9, 5, 64)[:, ::2, ::3, ::2]) is None


def test_header_is_plain_c_and_ctypes_mirrors_its_structs(tmp_path):
    """include/isi_b200.h compiles as C99 (it is what a non-Python host binds), and the ctypes
    structures have the sizes and field offsets the C compiler gives the header's."""
    import subprocess
    src = tmp_path / "probe.c"
    fields = [name for name, _ in _lib.MelifParams._fields_]
    inv_fields = [name for name, _ in _lib.ImelifParams._fields_]
    prints = "\n".join(
        [f'  printf("{f} %zu\\n", offsetof(isi_melif_params, {f}));' for f in fields] +
        [f'  printf("inv.{f} %zu\\n", offsetof(isi_imelif_params, {f}));' for f in inv_fields])
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "isi_b200.h"\n'
        "int main(void) {\n"
        '  printf("melif %zu\\n", sizeof(isi_melif_params));\n'
        '  printf("rows %zu\\n", sizeof(isi_rows_layout));\n'
        '  printf("imelif %zu\\n", sizeof(isi_imelif_params));\n'
        f"{prints}\n  return 0;\n}}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}",
                    str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                       text=True).stdout.splitlines())
    assert int(out["melif"]) == ctypes.sizeof(_lib.MelifParams)
    assert int(out["rows"]) == ctypes.sizeof(_lib.RowsLayout)
    assert int(out["imelif"]) == ctypes.sizeof(_lib.ImelifParams)
    for f in fields:
        assert int(out[f]) == getattr(_lib.MelifParams, f).offset, f
    for f in inv_fields:
        assert int(out["inv." + f]) == getattr(_lib.ImelifParams, f).offset, f
