"""Small, fixed workloads for `ncu` captures of single kernels (GPU box only):

    ncu --set full ... -k regex:vq_assign_pair_kernel -s 3 -c 1 python tools/ncu_targets.py pair
    ncu --set full ... -k regex:vq_assign_pstream     -s 3 -c 1 python tools/ncu_targets.py pstream

`pair`: K=512, D=64, 1 Mi rows (the quantiser roofline point of bench.py);
`pstream`: K=4096, D=128, 1 Mi rows (BASELINE config 4)."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.utils import synthetic  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "pair"
dim, n_embed = (64, 512) if what == "pair" else (128, 4096)
dev = torch.device("cuda:0")
embed = synthetic.synthetic_codebook(dim, n_embed)
m = QuantizedBottleneck(dim, n_embed).to(dev).eval()
m.embed.copy_(embed)
x = synthetic.synthetic_features(1 << 20, embed, 5).to(dev)
for _ in range(5):
    ind = m.assign(x)
torch.cuda.synchronize()
print(what, "ok", int(ind.max()))
