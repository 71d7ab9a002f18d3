"""TEST INFRASTRUCTURE ONLY -- near-tie accounting for code-map parity.

BASELINE.json's gate: code indices must match the reference bit-exactly at every position
whose reference distance gap exceeds 1e-5 relative; the near-tie count is reported.  Two code
maps are therefore never compared with a bare agreement percentage: every differing position
must be *explained* by the FP64 distances on the checker's own features.

A position where the product chose code ``g`` and the checker's FP64 nearest code is ``w`` is
accepted iff

    d64(x, g) - d64(x, w)  <=  rel * |d64(x, w)|  +  2 * safety * ||dx|| * ||e_g - e_w||

with ``x`` the checker's feature row, ``dx`` the measured difference between the product's
and the checker's feature rows (0 when both searched the very same features).  The second
term is exact first-order algebra: d_g(x+dx) - d_w(x+dx) = d_g(x) - d_w(x) - 2 dx.(e_g - e_w).

Reference: ``bottleneck.py:55-61`` (distance + argmax of -dist), ``vqvae.py:251-278`` (encode).
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class TieReport:
    n: int                 # positions compared
    differing: int         # product code != checker's FP32 code
    near_ties: int         # positions whose FP64 runner-up gap is <= rel (the reported count)
    differing_in_near_ties: int
    unexplained: int       # differing positions outside the bound: must be 0
    worst_excess: float    # max over differing positions of (d_g - d_w) / bound

    def __str__(self):
        return (f"{self.n} positions: {self.differing} differ, {self.near_ties} near ties "
                f"({self.differing_in_near_ties} of the differing ones), {self.unexplained} unexplained, "
                f"worst (d_got - d_best) / bound = {self.worst_excess:.3g}")


def explain_differences(features: torch.Tensor, got: torch.Tensor, want: torch.Tensor,
                        embed: torch.Tensor, feature_delta: Optional[torch.Tensor] = None,
                        rel: float = 1e-5, safety: float = 2.0) -> TieReport:
    """features [N, D] (the checker's), got / want [N] codes, embed [D, K] (bottleneck.py:49),
    feature_delta [N, D] = product features - checker features (or None)."""
    x = features.reshape(-1, features.shape[-1]).double().cpu()
    e = embed.double().cpu()
    got = got.reshape(-1).cpu().long()
    want = want.reshape(-1).cpu().long()
    # bottleneck.py:56-60 in FP64
    d = (x * x).sum(1, keepdim=True) - 2.0 * x @ e + (e * e).sum(0, keepdim=True)
    two = torch.topk(d, k=2, dim=1, largest=False)
    best, d1, d2 = two.indices[:, 0], two.values[:, 0], two.values[:, 1]
    gap = (d2 - d1) / d1.abs().clamp_min(1e-300)
    rows = torch.arange(x.shape[0])
    excess = d[rows, got] - d1                                        # >= 0
    bound = rel * d1.abs()
    if feature_delta is not None:
        dx = feature_delta.reshape(-1, x.shape[1]).double().cpu().norm(dim=1)
        bound = bound + 2.0 * safety * dx * (e[:, got] - e[:, best]).norm(dim=0)
    differing = got != want
    bad = differing & (excess > bound)
    ratio = (excess / bound.clamp_min(1e-300))[differing]
    return TieReport(n=int(x.shape[0]), differing=int(differing.sum()), near_ties=int((gap <= rel).sum()),
                     differing_in_near_ties=int((differing & (gap <= rel)).sum()),
                     unexplained=int(bad.sum()), worst_excess=float(ratio.max()) if ratio.numel() else 0.0)


def encode_with_features(model, spec):
    """``VQVAE.encode`` (vqvae.py:251-278) step by step on any module with the reference's
    attribute names (the unmodified reference class or this repo's), returning the pre-quantiser
    feature rows next to the codes: (feat_t [B,Ht,Wt,D], id_t, feat_b [B,Hb,Wb,D], id_b)."""
    enc_b = model.enc_b(spec)
    enc_t = model.enc_t(enc_b)
    feat_t = model.quantize_conv_t(enc_t).permute(0, 2, 3, 1)
    quant_t, _, id_t, _ = model.quantize_t(feat_t)
    dec_t = model.dec_t(quant_t.permute(0, 3, 1, 2))
    feat_b = model.quantize_conv_b(torch.cat([dec_t, enc_b], 1)).permute(0, 2, 3, 1)
    _, _, id_b, _ = model.quantize_b(feat_b)
    return feat_t, id_t, feat_b, id_b


def reference_or_port_model(state_dict, model_kw):
    """The CPU checker of the encoder: the UNMODIFIED reference ``VQVAE`` (from /root/reference or
    the staged baseline/_ref copy) when it can be imported, else this repo's wiring with the
    oracle quantiser.  Returns (model.eval(), kind) with kind in {"reference", "port"}."""
    from oracle import ref_loader
    if ref_loader.available():
        cls = ref_loader.load_reference_vqvae_class()
        model = cls(**model_kw).eval()
        missing = model.load_state_dict(state_dict, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        return model, "reference"
    from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE
    from oracle import quantizer_oracle as qo
    model = VQVAE(**model_kw, bottleneck_cls=qo.OracleBottleneck).eval()
    model.load_state_dict(state_dict)
    return model, "port"


def explain_code_maps(checker, product, got_t, got_b, embed_t, embed_b):
    """Top and bottom reports for code maps ``got_t`` / ``got_b`` against a checker pass.

    ``checker`` and ``product`` are ``encode_with_features`` 4-tuples: the checker's features and
    codes, and the product path's own features (their difference enters the bound).  The bottom
    level is compared on the notes whose top maps agree everywhere (checker, product features'
    pass and ``got_t``): bottom features depend on the top codes through ``dec_t``
    (vqvae.py:264-270), so a legitimate top near-tie flip changes everything below it.
    Returns (top report, bottom report, number of notes compared at the bottom level)."""
    feat_t, want_t, feat_b, want_b = [v.detach().cpu() for v in checker]
    pfeat_t, pid_t, pfeat_b, _ = [v.detach().cpu() for v in product]
    got_t, got_b = got_t.detach().cpu(), got_b.detach().cpu()
    n = want_t.shape[0]
    rep_t = explain_differences(feat_t, got_t, want_t, embed_t, pfeat_t - feat_t)
    same_top = ((got_t == want_t) & (pid_t == want_t)).reshape(n, -1).all(1)
    rep_b = explain_differences(feat_b[same_top], got_b[same_top], want_b[same_top], embed_b,
                                (pfeat_b - feat_b)[same_top])
    return rep_t, rep_b, int(same_top.sum())
