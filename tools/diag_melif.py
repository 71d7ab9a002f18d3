"""GPU diagnostic: the warp-specialised and the generic front-end kernels on the same input
(ISI_MELIF_GENERIC=1 forces the generic one; run each in its own process)."""
import os
import subprocess
import sys
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if len(sys.argv) > 1:
    import torch
    from interactive_spectrogram_inpainting_b200.utils import synthetic
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    outs = {}
    for hop, n in ((512, 64000), (500, 40000)):
        h = MelSpectrogramsHelper(hop_length=hop).to("cuda:0")
        a = synthetic.synthetic_notes(3, n_samples=n)
        pcm = (a * 32767).round().to(torch.int16)
        outs[f"f32_{hop}"] = h.to_spectrogram((pcm.float() * h.pcm_scale).to("cuda:0")).cpu()
        outs[f"pcm_{hop}"] = h.to_spectrogram(pcm.to("cuda:0")).cpu()
        outs[f"f32_{hop}_again"] = h.to_spectrogram((pcm.float() * h.pcm_scale).to("cuda:0")).cpu()
    torch.save(outs, sys.argv[1])
    sys.exit(0)

import torch
for tag, env in (("ws", {}), ("generic", {"ISI_MELIF_GENERIC": "1"})):
    subprocess.run([sys.executable, __file__, f"/tmp/diag_{tag}.pt"], env={**os.environ, **env}, check=True)
ws, gen = torch.load("/tmp/diag_ws.pt"), torch.load("/tmp/diag_generic.pt")
def cmp(name, a, b):
    d = (a - b).abs()
    idx = (d > 0).nonzero()
    print(f"{name}: equal={torch.equal(a, b)} max|diff|={float(d.max()):.3g} n_diff={len(idx)} of {a.numel()}",
          "first:", idx[:6].tolist(), "frames:", sorted(set(idx[:, 3].tolist()))[:20] if len(idx) else "")
for k in ws:
    cmp(f"ws vs generic [{k}]", ws[k], gen[k])
for hop in (512, 500):
    cmp(f"ws: f32 vs pcm [{hop}]", ws[f"f32_{hop}"], ws[f"pcm_{hop}"])
    cmp(f"ws: run-to-run [{hop}]", ws[f"f32_{hop}"], ws[f"f32_{hop}_again"])
    cmp(f"generic: f32 vs pcm [{hop}]", gen[f"f32_{hop}"], gen[f"pcm_{hop}"])
    cmp(f"generic: run-to-run [{hop}]", gen[f"f32_{hop}"], gen[f"f32_{hop}_again"])
