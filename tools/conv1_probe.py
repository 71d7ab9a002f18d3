"""First encoder conv (2 -> 32 channels, k4 s2 p1) at the bench batch: does padding the input
channels give cuDNN a better kernel?"""
import json
import torch

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
B = 444
out = {}
for cin in (2, 4, 8):
    x = torch.randn(B, cin, 1024, 128, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(32, cin, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(32, device=dev)
    for name, fn in (("fused", lambda: torch.cudnn_convolution_relu(x, w, b, (2, 2), (1, 1), (1, 1), 1)),
                     ("plain", lambda: torch.nn.functional.conv2d(x, w, None, 2, 1))):
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
        out[f"cin{cin}_{name}"] = round(a.elapsed_time(e) / 10, 3)
print(json.dumps(out))
