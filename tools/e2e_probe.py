"""Where the end-to-end extraction step spends its time (host clock, one B200)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from interactive_spectrogram_inpainting_b200 import extract  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE  # noqa: E402

B, K = 444, 10
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
helper = MelSpectrogramsHelper(channels_last=True).to(dev)
model = VQVAE(**bench.MODEL_KW).to(dev).eval().to(memory_format=torch.channels_last)
host_audio = bench.make_audio(B).pin_memory()
names = [f"n{i}" for i in range(B)]
out = {}


def timed(label, fn, reps=1):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    out[label] = round((time.perf_counter() - t) * 1e3 / reps / K, 3)


def h2d_only():
    for _ in range(K):
        host_audio.to(dev, non_blocking=True)


def loader_only():
    for spec, _ in extract.SpectrogramBatches([(host_audio, names)] * K, helper, dev):
        pass


def loader_noprefetch():
    for spec, _ in extract.SpectrogramBatches([(host_audio, names)] * K, helper, dev, prefetch=False):
        pass


def full():
    extract.extract_codes(extract.SpectrogramBatches([(host_audio, names)] * K, helper, dev), model)


def full_noprefetch():
    extract.extract_codes(extract.SpectrogramBatches([(host_audio, names)] * K, helper, dev, prefetch=False), model)


def resident():
    a = host_audio.to(dev)
    with torch.no_grad():
        for _ in range(K):
            model.encode_codes(helper.to_spectrogram(a))


for label, fn in (("h2d_only", h2d_only), ("loader_only", loader_only), ("loader_noprefetch", loader_noprefetch),
                  ("resident_compute", resident), ("full", full), ("full_noprefetch", full_noprefetch), ("full_again", full)):
    timed(label, fn, reps=2)
print(json.dumps({"ms_per_step": out}))
