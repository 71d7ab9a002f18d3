"""Per-batch timeline of the end-to-end extraction loop (host clock + CUDA events)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE  # noqa: E402

B, K = 444, int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
helper = MelSpectrogramsHelper(channels_last=True).to(dev)
model = VQVAE(**bench.MODEL_KW).to(dev).eval().to(memory_format=torch.channels_last)
host_audio = bench.make_audio(B).pin_memory()
main = torch.cuda.current_stream(dev)
side = torch.cuda.Stream(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
host_t = torch.empty(B, 32, 4, dtype=torch.int64).pin_memory()
host_b = torch.empty(B, 64, 8, dtype=torch.int64).pin_memory()


def upload():
    with torch.cuda.stream(side):
        a, b = ev(), ev()
        a.record(side)
        d = host_audio.to(dev, non_blocking=True)
        b.record(side)
    return d, a, b


log = []
with torch.no_grad():
    for _ in range(3):
        model.encode_codes(helper.to_spectrogram(host_audio.to(dev)))
    torch.cuda.synchronize()
    origin = ev(); origin.record()
    t_origin = time.perf_counter()
    pending = upload()
    prev_done = None
    for i in range(K):
        d, ha, hb = pending
        pending = upload() if i + 1 < K else None
        t0 = time.perf_counter()
        main.wait_event(hb)
        d.record_stream(main)
        c0, c1, c2 = ev(), ev(), ev()
        c0.record()
        spec = helper.to_spectrogram(d)
        c1.record()
        id_t, id_b = model.encode_codes(spec)
        host_t.copy_(id_t, non_blocking=True)
        host_b.copy_(id_b, non_blocking=True)
        c2.record()
        t1 = time.perf_counter()
        if prev_done is not None:
            prev_done.synchronize()
        t2 = time.perf_counter()
        prev_done = c2
        log.append((i, ha, hb, c0, c1, c2, t0 - t_origin, t1 - t_origin, t2 - t_origin))
    torch.cuda.synchronize()
    total = time.perf_counter() - t_origin
rows = []
for i, ha, hb, c0, c1, c2, t0, t1, t2 in log:
    rows.append({"i": i, "h2d": [round(origin.elapsed_time(ha), 2), round(origin.elapsed_time(hb), 2)],
                 "melif": [round(origin.elapsed_time(c0), 2), round(origin.elapsed_time(c1), 2)],
                 "encode_end": round(origin.elapsed_time(c2), 2),
                 "host_launch": [round(t0 * 1e3, 2), round(t1 * 1e3, 2)], "host_after_sync": round(t2 * 1e3, 2)})
for r in rows:
    print(json.dumps(r))
print("ms_per_step", round(total * 1e3 / K, 3), "mem GB", torch.cuda.max_memory_allocated() / 1e9,
      "reserved", torch.cuda.memory_reserved() / 1e9)
