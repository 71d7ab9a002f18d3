import json, sys, time
import torch
sys.path.insert(0, ".")
import bench
from interactive_spectrogram_inpainting_b200 import extract
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE

B, K = 444, 20
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
helper = MelSpectrogramsHelper(channels_last=True).to(dev)
model = VQVAE(**bench.MODEL_KW).to(dev).eval().to(memory_format=torch.channels_last)
host_audio = bench.make_audio(B).pin_memory()
names = [f"n{i}" for i in range(B)]
a = host_audio.to(dev)
with torch.no_grad():
    for _ in range(23):
        model.encode_codes(helper.to_spectrogram(a))
torch.cuda.synchronize()
def stats():
    s = torch.cuda.memory_stats()
    return {"dev_alloc": s["num_device_alloc"], "dev_free": s["num_device_free"], "reserved_GB": round(s["reserved_bytes.all.current"]/1e9,2), "retries": s["num_alloc_retries"]}
print("after resident", stats())
for run in range(4):
    stamps = []
    t = time.perf_counter()
    extract.extract_codes(extract.SpectrogramBatches([(host_audio, names)] * K, helper, dev), model,
                          sink=lambda rows: stamps.append(time.perf_counter()))
    torch.cuda.synchronize()
    total = time.perf_counter() - t
    gaps = [round((b - a) * 1e3, 2) for a, b in zip(stamps[:-1], stamps[1:])]
    print("run", run, "ms/step", round(total * 1e3 / K, 2), "gaps", gaps, stats())
