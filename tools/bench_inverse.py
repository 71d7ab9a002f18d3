#!/usr/bin/env python
"""Secondary measurement for the inverse front end (not the driver's bench
line): isi_melif_inverse (``to_audio``) at the server's batch sizes and at the extraction
batch, next to the same arithmetic as plain torch ops on the same GPU.

Prints one JSON object; run on the GPU box: python tools/bench_inverse.py > gpurun_out/inverse.json
"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper  # noqa: E402

DEV = torch.device("cuda:0")
PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}


def timed(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                       # evict the inputs from the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters


def main():
    helper = MelSpectrogramsHelper().to(DEV)
    g = torch.Generator().manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    rows = []
    only = int(sys.argv[sys.argv.index("--only") + 1]) if "--only" in sys.argv else None
    for b in ((only,) if only else (1, 2, 4, 8, 16, 444)):
        spec = torch.stack([torch.randn(b, 1024, 128, generator=g) * 2 - 3,
                            torch.rand(b, 1024, 128, generator=g) * 2 - 1], 1).to(DEV)
        with torch.no_grad():
            ms = timed(lambda: helper.to_audio(spec), flush=flush if b >= 16 else None)
            torch_ms = None if (only or "--kernel-only" in sys.argv) else timed(lambda: helper.to_audio_differentiable(spec), iters=5, warmup=2)
        bytes_alg = b * (2 * 1024 * 128 * 4 + 64000 * 4)
        rows.append({"batch": b, "kernel_ms": ms, "torch_ops_ms": torch_ms,
                     "us_per_note": 1e3 * ms / b, "algorithmic_GBps": bytes_alg / (ms * 1e-3) / 1e9})
    hbm = PEAKS.get("hbm_gbs")
    print(json.dumps({"inverse_front_end": rows, "algorithmic_bytes_per_note": 2 * 1024 * 128 * 4 + 64000 * 4,
                      "hbm_peak_GBps": hbm}))


if __name__ == "__main__":
    main()
