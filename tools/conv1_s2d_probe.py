"""First encoder conv as a 3x3 stride-1 conv on the space-to-depth spectrogram (8 channels)."""
import json
import torch

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
B = 444
out = {}


def timed(fn):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
    return round(a.elapsed_time(e) / 10, 3)


x = torch.randn(B, 2, 1024, 128, device=dev).contiguous(memory_format=torch.channels_last)
w = torch.randn(32, 2, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)
b = torch.randn(32, device=dev)
out["k4s2_cin2"] = timed(lambda: torch.cudnn_convolution_relu(x, w, b, (2, 2), (1, 1), (1, 1), 1))
for cin, k in ((8, 3), (8, 2)):
    xs = torch.randn(B, cin, 512, 64, device=dev).contiguous(memory_format=torch.channels_last)
    ws = torch.randn(32, cin, k, k, device=dev).contiguous(memory_format=torch.channels_last)
    pad = (1, 1) if k == 3 else (0, 0)
    out[f"k{k}s1_cin{cin}"] = timed(lambda: torch.cudnn_convolution_relu(xs, ws, b, (1, 1), pad, (1, 1), 1))
# unfold-free alternative: shift the s2d grid by one pixel so a 2x2 stride-1 conv suffices
xs = torch.randn(B, 8, 513, 65, device=dev).contiguous(memory_format=torch.channels_last)
ws = torch.randn(32, 8, 2, 2, device=dev).contiguous(memory_format=torch.channels_last)
out["k2s1_cin8_shifted_grid"] = timed(lambda: torch.cudnn_convolution_relu(xs, ws, b, (1, 1), (0, 0), (1, 1), 1))
print(json.dumps(out))
