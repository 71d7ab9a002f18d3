"""GPU: time the front-end kernel alone (444 NSynth-shaped notes, 16-bit PCM in, space-to-depth
blocks out -- the extraction path's call) under sets of environment knobs, one process each.

    python tools/time_melif.py                       # the default build
    python tools/time_melif.py ISI_MELIF_WAIT_HINT=0,ISI_MELIF_WAIT_SLEEP=0 ISI_MELIF_WAIT_SLEEP=0
"""
import os
import subprocess
import sys
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if os.environ.get("_TIME_MELIF_CHILD"):
    import torch
    from interactive_spectrogram_inpainting_b200.utils import synthetic
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    dev = torch.device("cuda:0")
    mode = os.environ.get("TIME_MELIF_LAYOUT", "s2d")
    helper = MelSpectrogramsHelper(**{"s2d": dict(space_to_depth=True), "s2dt": dict(space_to_depth="transposed"),
                                      "planar": {}, "cl": dict(channels_last=True)}[mode]).to(dev)
    pcm = (synthetic.synthetic_notes(12) * 32767).round().to(torch.int16).repeat(37, 1).to(dev)   # 444 notes
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        out = helper.to_spectrogram(pcm)
    times = []
    for _ in range(40):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); out = helper.to_spectrogram(pcm); t1.record()
        torch.cuda.synchronize()
        times.append(t0.elapsed_time(t1))
    times.sort()
    print(f"median {times[len(times) // 2]:.4f} ms  min {times[0]:.4f}  p90 {times[int(len(times) * .9)]:.4f}"
          f"  checksum {float(out.double().sum()):.6f}")
    sys.exit(0)

configs = sys.argv[1:] or [""]
for cfg in configs:
    env = {**os.environ, "_TIME_MELIF_CHILD": "1"}
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        env[k] = v
    r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
    print(f"[{cfg or 'default'}] {r.stdout.strip() or r.stderr.strip()[-400:]}", flush=True)
