#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3-5 (not the driver's bench line):

  cfg 3  quantiser part of a training step at the reference batch (64 notes/GPU: 8 192 top +
         32 768 bottom rows): forward + EMA statistics + EMA update, per step
  cfg 4  large-codebook stress: K=4096, D=128 nearest-code sweep over N
  cfg 5  interactive decode path: embed_code of edited top/bottom code maps (+ the torch conv
         decoder) at batch 1..16, eager and replayed from a CUDA graph

Prints one JSON object; run on the GPU box: python tools/bench_extra.py > gpurun_out/extra.json
"""
import json
import pathlib
import statistics
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from interactive_spectrogram_inpainting_b200.utils import synthetic  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE, GraphedDecodeCode  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def cfg3():
    out = {}
    embed = synthetic.synthetic_codebook(64, 512)
    mods = {}
    for name, rows in (("top", 64 * 32 * 4), ("bottom", 64 * 64 * 8)):
        m = QuantizedBottleneck(64, 512).to(DEV).train()
        m.embed.copy_(embed); m.embed_avg.copy_(embed)
        x = synthetic.synthetic_features(rows, embed, 11).to(DEV).requires_grad_(True)
        mods[name] = (m, x)

    def step():
        for m, x in mods.values():
            q, diff, ind, perp = m(x)
            (q.sum() * 1e-3 + 0.25 * diff).backward()
            x.grad = None
    out["ms_per_step_fwd_bwd_ema_top_plus_bottom"] = timed(step)
    out["rows_per_step"] = 64 * (128 + 512)
    return out


def cfg4():
    out = []
    for dim, n_embed, algos in ((128, 4096, ("auto", "simt")), (64, 4096, ("auto", "tcgen05", "simt"))):
        embed = synthetic.synthetic_codebook(dim, n_embed)
        m = QuantizedBottleneck(dim, n_embed).to(DEV).eval()
        m.embed.copy_(embed)
        for n in (8192, 65536, 262144, 1048576):
            x = synthetic.synthetic_features(n, embed, 5).to(DEV)
            for algo in algos:
                if algo == "simt" and n > 262144:
                    continue
                m.assign_algo = algo
                ms = timed(lambda: m.assign(x), iters=5, warmup=2)
                out.append({"dim": dim, "n_embed": n_embed, "rows": n, "algo": algo, "ms": ms,
                            "algorithmic_tflops": 2.0 * n * n_embed * dim / (ms * 1e-3) / 1e12})
    return out


def cfg5():
    from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper
    helper = MelSpectrogramsHelper().to(DEV)
    out = []
    torch.manual_seed(0)
    model = VQVAE(in_channel=2, resolution_factors={'bottom': 16, 'top': 2},
                  adapt_quantized_durations=False).to(DEV).eval()
    for b in (1, 2, 4, 8, 16):
        top, bottom = synthetic.synthetic_codemaps(b)
        top, bottom = top.to(DEV), bottom.to(DEV)
        with torch.no_grad():
            lookup_ms = timed(lambda: (model.quantize_t.embed_code(top), model.quantize_b.embed_code(bottom)))
            eager_ms = timed(lambda: model.decode_code(top, bottom))
            # CUDA graph of the whole decode (lookup kernels + conv decoder)
            graphed = GraphedDecodeCode(model, top, bottom)
            graph_ms = timed(lambda: graphed(top, bottom))
            assert torch.allclose(graphed(top, bottom), model.decode_code(top, bottom), rtol=1e-3, atol=1e-4)
            # the server's whole request: codes -> spectrogram -> audio (flask_server.py:593-596)
            to_audio_ms = timed(lambda: helper.to_audio(graphed(top, bottom)))
            to_audio_torch_ms = timed(lambda: helper.to_audio_differentiable(graphed(top, bottom)), iters=5)
            full = GraphedDecodeCode(model, top, bottom, to_audio=helper)
            full_ms = timed(lambda: full(top, bottom))
        out.append({"batch": b, "embed_code_top_plus_bottom_ms": lookup_ms,
                    "decode_code_eager_ms": eager_ms, "decode_code_cuda_graph_ms": graph_ms,
                    "codes_to_audio_graph_plus_kernel_ms": to_audio_ms,
                    "codes_to_audio_graph_plus_torch_inverse_ms": to_audio_torch_ms,
                    "codes_to_audio_one_cuda_graph_ms": full_ms})
    return out


if __name__ == "__main__":
    print(json.dumps({"cfg3_train_quantiser": cfg3(), "cfg4_large_codebook": cfg4(),
                      "cfg5_decode_latency": cfg5()}, indent=1))
