"""Host->device copy rate of this box for the bench's per-step audio batch (pinned memory)."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for name, shape, dtype in (("f32_444_notes", (444, 64000), torch.float32),
                           ("i16_444_notes", (444, 64000), torch.int16),
                           ("f32_1GiB", (1 << 28,), torch.float32)):
    host = torch.empty(shape, dtype=dtype).pin_memory()
    devbuf = torch.empty(shape, dtype=dtype, device=dev)
    for _ in range(3):
        devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        devbuf.copy_(host, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    out[name] = {"ms": round(ms, 3), "GB/s": round(host.numel() * host.element_size() / ms / 1e6, 2)}
    a.record()
    for _ in range(10):
        host.copy_(devbuf, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out[name]["d2h_GB/s"] = round(host.numel() * host.element_size() / (a.elapsed_time(b) / 10) / 1e6, 2)
print(json.dumps(out))
