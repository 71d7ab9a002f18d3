"""GPU-box helper: time isi_vq_gather_stats (training mode) over row counts."""
import pathlib, sys, statistics
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from interactive_spectrogram_inpainting_b200 import _lib
from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
dev = torch.device("cuda:0")
m = QuantizedBottleneck(64, 512).to(dev).train(); m.sync_ema_stats = False
for n in (65536, 262144, 1048576, 4194304):
    for kind in ("shuffled", "uniform-random-codes", "single-code"):
        if kind == "shuffled":
            x = synthetic.synthetic_features(n, m.embed.cpu()).to(dev)
        elif kind == "uniform-random-codes":
            pick = torch.randint(0, 512, (n,), device=dev)
            x = (m.embed.t()[pick] + 0.01 * torch.randn(n, 64, device=dev)).contiguous()
        else:
            x = (m.embed[:, 7][None, :] + 0.01 * torch.randn(n, 64, device=dev)).contiguous()
        emb0 = m.embed.clone()
        for _ in range(2): m(x)
        torch.cuda.synchronize(); _lib.event_log = []
        for _ in range(5):
            m.embed.copy_(emb0)
            m(x)
        torch.cuda.synchronize(); ev, _lib.event_log = _lib.event_log, None
        t = statistics.mean(a.elapsed_time(b) for name, a, b in ev if name == "isi_vq_gather_stats")
        print(f"rows {n:8d} {kind:22s} gather_stats {t*1e3:8.1f} us  {n*520/t/1e6:7.1f} GB/s")
