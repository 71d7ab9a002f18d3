#!/usr/bin/env python
"""Secondary measurement for the front end (not the driver's bench line): isi_melif_forward
(``to_spectrogram``) at the server's batch sizes and at extraction batches, in every output
layout, FP32 and int16 PCM input.

Prints one JSON object; run on the GPU box: python tools/bench_forward.py > gpurun_out/forward.json
"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.utils import synthetic  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper  # noqa: E402

DEV = torch.device("cuda:0")
PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
LAYOUTS = {"planar": {}, "channels_last": dict(channels_last=True), "blocks": dict(space_to_depth=True),
           "blocks_frequency_fastest": dict(space_to_depth="transposed")}


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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    base = synthetic.synthetic_notes(16)
    rows = []
    for name, kw in LAYOUTS.items():
        helper = MelSpectrogramsHelper(**kw).to(DEV)
        for b in (1, 4, 16, 148, 444, 888):
            audio = base.repeat((b + 15) // 16, 1)[:b]
            for fmt in ("pcm16", "f32"):
                x = ((audio * 32767).round().to(torch.int16) if fmt == "pcm16" else audio).to(DEV)
                ms = timed(lambda: helper.to_spectrogram(x), flush=flush if b >= 16 else None)
                bytes_alg = b * (2 * 1024 * 128 * 4 + 64000 * x.element_size())
                rows.append({"layout": name, "audio": fmt, "batch": b, "kernel_ms": round(ms, 5),
                             "us_per_note": round(1e3 * ms / b, 3),
                             "algorithmic_GBps": round(bytes_alg / (ms * 1e-3) / 1e9, 1)})
    print(json.dumps({"front_end": rows, "hbm_peak_GBps": PEAKS.get("hbm_gbs")}))


if __name__ == "__main__":
    main()
