"""One small call of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.utils import synthetic  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import (  # noqa: E402
    MelSpectrogramsHelper, SpectrogramsHelper)
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "frontend"):
    audio = synthetic.synthetic_notes(2, n_samples=16000).to(dev)
    pcm = (audio * 32767).round().to(torch.int16)
    for helper in (MelSpectrogramsHelper(), MelSpectrogramsHelper(space_to_depth=True, n_frames=36),
                   MelSpectrogramsHelper(space_to_depth="transposed", n_frames=36),
                   MelSpectrogramsHelper(channels_last=True, n_frames=32), MelSpectrogramsHelper(n_frames=35),
                   SpectrogramsHelper(channels_last=True), MelSpectrogramsHelper(n_fft=512, hop_length=125, window_length=512)):
        helper = helper.to(dev)
        helper.to_spectrogram(audio)
        helper.to_spectrogram(pcm)
    torch.cuda.synchronize()
    print("frontend ok")

if which in ("all", "inverse"):
    g = torch.Generator().manual_seed(0)
    for helper, frames, seg in ((MelSpectrogramsHelper(), 32, 8), (SpectrogramsHelper(), 12, None),
                                (MelSpectrogramsHelper(n_fft=512, hop_length=125, window_length=512), 13, 8),
                                (MelSpectrogramsHelper(n_fft=1024, hop_length=256, window_length=1024), 18, None)):
        helper = helper.to(dev)
        helper.inverse_seg_frames = seg
        spec = torch.stack([torch.randn(2, helper.n_freq, frames, generator=g) - 3,
                            torch.rand(2, helper.n_freq, frames, generator=g) * 2 - 1], 1).to(dev)
        helper.to_audio(spec)
    torch.cuda.synchronize()
    print("inverse ok")

if which in ("all", "projection"):
    from interactive_spectrogram_inpainting_b200.vqvae.vqvae import PointwiseProjection
    torch.manual_seed(0)
    for c0, c1, shape in ((128, 0, (3, 32, 4)), (64, 128, (5, 37, 11)), (192, 64, (2, 150, 9))):
        conv = torch.nn.Conv2d(c0 + c1, 64, 1).to(dev)
        srcs = [torch.randn(shape[0], c, *shape[1:], device=dev).contiguous(memory_format=torch.channels_last)
                for c in (c0, c1) if c]
        proj = PointwiseProjection(conv)
        proj.min_rows = 1
        with torch.no_grad():
            assert proj.usable(srcs)
            proj(srcs)
    torch.cuda.synchronize()
    print("projection ok")

if which in ("all", "quantizer"):
    embed = synthetic.synthetic_codebook(64, 512)
    for algo, rows, k in (("simt", 700, 512), ("tcgen05", 5000, 512), ("tcgen05_pair", 5000, 512),
                          ("tcgen05_pair_stream", 5000, 1024)):
        embed = synthetic.synthetic_codebook(64, k)
        m = QuantizedBottleneck(64, k).to(dev).eval()
        m.embed.copy_(embed)
        m.assign_algo = algo
        x = synthetic.synthetic_features(rows, embed).to(dev)
        q, d, ind, p = m(x)
        m.embed_code(ind.view(10, -1, 10)[:, :7].contiguous())
    embed = synthetic.synthetic_codebook(64, 512)
    m = QuantizedBottleneck(64, 512).to(dev).train()
    m.embed.copy_(embed); m.embed_avg.copy_(embed)
    x = synthetic.synthetic_features(64 * 148 * 3 + 17, embed).to(dev)
    for _ in range(2):
        m(x)                                             # smem statistics kernel + EMA update
    m(synthetic.synthetic_features(900, embed).to(dev).view(100, 9, 64).permute(1, 0, 2))   # generic (strided) kernel
    big = QuantizedBottleneck(128, 4096).to(dev).eval()
    big(torch.randn(4500, 128, device=dev))              # streaming pair kernel, D = 128
    torch.cuda.synchronize()
    print("quantizer ok")
