"""TEST INFRASTRUCTURE ONLY -- regenerates ``tests/golden/*.npz``.

Run in the dev container, where the read-only reference checkout exists:

    python -m oracle.make_golden

Quantiser fixtures come from the UNMODIFIED reference class
(``interactive_spectrogram_inpainting/vqvae/bottleneck.py:30-104``, imported by
``oracle/ref_loader.py``) on seeded inputs.  The front-end fixture comes from
``oracle/frontend_oracle.py`` itself (the reference ships no front-end source:
parity unpinned) and only guards that restatement against accidental change.
"""
import pathlib
import warnings

import numpy as np
import torch

from interactive_spectrogram_inpainting_b200.utils import synthetic
from oracle import frontend_oracle, ref_loader

GOLDEN = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"


def _ref_module(dim, n_embed, embed, **kw):
    ref = ref_loader.load_reference_bottleneck()
    q = ref.QuantizedBottleneck(dim, n_embed, **kw)
    q.embed.copy_(embed)
    q.embed_avg.copy_(embed)
    return q


def quantizer_eval_fixture():
    """cfg 1 shapes (top 32x4, bottom 64x8, K=512, D=64) at batch 2, eval mode."""
    embed = synthetic.synthetic_codebook(64, 512)
    out = {"embed": embed.numpy()}
    for name, hw, seed in (("top", (32, 4), 1234), ("bottom", (64, 8), 1235)):
        x = synthetic.synthetic_features(2 * hw[0] * hw[1], embed, seed).view(2, *hw, 64)
        q = _ref_module(64, 512, embed).eval()
        with torch.no_grad():
            quant, diff, ind, perp = q(x)
        out.update({f"x_{name}": x.numpy(), f"quantize_{name}": quant.numpy(),
                    f"diff_{name}": diff.numpy(), f"ind_{name}": ind.numpy(),
                    f"perplexity_{name}": perp.numpy()})
    np.savez_compressed(GOLDEN / "quantizer_eval_cfg1.npz", **out)


def quantizer_train_fixture():
    """three EMA steps (bottleneck.py:79-92) on 640-row batches, K=512, D=64."""
    embed = synthetic.synthetic_codebook(64, 512)
    q = _ref_module(64, 512, embed).train()
    out = {"embed0": embed.numpy()}
    for step in range(3):
        x = synthetic.synthetic_features(640, embed, synthetic.FEATURE_SEED_TRAIN + step)
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            quant, diff, ind, perp = q(x)
        out.update({f"x{step}": x.numpy(), f"ind{step}": ind.numpy(),
                    f"quantize{step}": quant.numpy(), f"diff{step}": diff.numpy(),
                    f"perplexity{step}": perp.numpy(),
                    f"embed_after{step}": q.embed.numpy().copy(),
                    f"cluster_size_after{step}": q.cluster_size.numpy().copy(),
                    f"embed_avg_after{step}": q.embed_avg.numpy().copy()})
    np.savez_compressed(GOLDEN / "quantizer_train_3steps.npz", **out)


def quantizer_edge_fixture():
    """ragged shapes, exact ties (duplicated codewords -> lowest index wins,
    bottleneck.py:61), a permuted NCHW view as input (vqvae.py:260)."""
    g = torch.Generator().manual_seed(99)
    embed = torch.randn(8, 20, generator=g)
    embed[:, 7] = embed[:, 3]            # exact duplicate: index 3 must win
    embed[:, 19] = embed[:, 0]
    nchw = torch.randn(3, 8, 5, 7, generator=g)
    nchw[0, :, 0, 0] = embed[:, 3]       # zero distance to codes 3 and 7
    nchw[1, :, 2, 2] = embed[:, 19]
    x = nchw.permute(0, 2, 3, 1)         # non-contiguous [3,5,7,8]
    q = _ref_module(8, 20, embed).eval()
    with torch.no_grad():
        quant, diff, ind, perp = q(x)
        codes = torch.randint(0, 20, (2, 3, 4), generator=g)
        looked_up = q.embed_code(codes)
    np.savez_compressed(GOLDEN / "quantizer_edge.npz", embed=embed.numpy(),
                        nchw=nchw.numpy(), quantize=quant.numpy(), diff=diff.numpy(),
                        ind=ind.numpy(), perplexity=perp.numpy(),
                        codes=codes.numpy(), looked_up=looked_up.numpy())


def frontend_fixture():
    """strided sample of the restated mel-IF and linear-IF spectrograms of 2 notes."""
    audio = synthetic.synthetic_notes(2)
    mel = frontend_oracle.to_spectrogram(audio, frontend_oracle.FrontEndConfig())
    lin = frontend_oracle.to_spectrogram(
        audio, frontend_oracle.FrontEndConfig(use_mel_scale=False))
    np.savez_compressed(GOLDEN / "frontend_unpinned.npz",
                        audio_head=audio[:, :4096].numpy(),
                        mel=mel[:, :, ::8, ::4].numpy(), lin=lin[:, :, ::8, ::4].numpy())


def inverse_fixture():
    """strided sample of the restated inverse (to_audio) of a seeded random mel / linear
    spectrogram: 24 frames -> 10 752 samples."""
    g = torch.Generator().manual_seed(20200117)
    spec = torch.stack([torch.randn(2, 1024, 24, generator=g) * 2.0 - 3.0,
                        torch.rand(2, 1024, 24, generator=g) * 2.0 - 1.0], 1)
    mel = frontend_oracle.to_audio(spec.double(), frontend_oracle.FrontEndConfig())
    lin = frontend_oracle.to_audio(spec.double(), frontend_oracle.FrontEndConfig(use_mel_scale=False))
    np.savez_compressed(GOLDEN / "inverse_unpinned.npz", spec_corner=spec[:, :, :8, :8].numpy(),
                        mel=mel[:, ::7].float().numpy(), lin=lin[:, ::7].float().numpy())


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)             # fixed reduction order for the fixtures
    quantizer_eval_fixture()
    quantizer_train_fixture()
    quantizer_edge_fixture()
    frontend_fixture()
    inverse_fixture()
    for p in sorted(GOLDEN.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
