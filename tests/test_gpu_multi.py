"""Two-GPU (NCCL) check of the training exchange step: the packed EMA statistics are
all-reduced so that both ranks update like one process on the concatenated batch, and end with
bit-identical codebooks.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interactive_spectrogram_inpainting_b200.utils import distributed as du
from interactive_spectrogram_inpainting_b200.utils import synthetic
from oracle import quantizer_oracle as qo

pytestmark = pytest.mark.gpu
DIM, K, ROWS, STEPS = 64, 512, 9000, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        embed = synthetic.synthetic_codebook(DIM, K)
        m = QuantizedBottleneck(DIM, K).to(f"cuda:{rank}").train()
        m.embed.copy_(embed); m.embed_avg.copy_(embed)
        inds = []
        for step in range(STEPS):
            full = synthetic.synthetic_features(ROWS, embed, 300 + step)
            lo, hi = du.shard_range(ROWS, rank, world)
            _, _, ind, _ = m(full[lo:hi].to(f"cuda:{rank}"))
            inds.append(ind.cpu())
        torch.save({"embed": m.embed.cpu(), "cluster_size": m.cluster_size.cpu(),
                    "embed_avg": m.embed_avg.cpu(), "inds": inds}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_allreduced_ema_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for key in ("embed", "cluster_size", "embed_avg"):
        assert torch.equal(got[0][key], got[1][key]), key          # ranks stay bit-identical
    embed = synthetic.synthetic_codebook(DIM, K)
    st = qo.CodebookState(embed.clone(), torch.zeros(K), embed.clone())
    for step in range(STEPS):
        full = synthetic.synthetic_features(ROWS, embed, 300 + step)
        ind = torch.cat([got[r]["inds"][step] for r in range(world)])   # the kernels' own codes
        qo.ema_update(st, full, ind, 0.99, 1e-5)
    for key, want in (("cluster_size", st.cluster_size), ("embed_avg", st.embed_avg), ("embed", st.embed)):
        err = (got[0][key] - want).abs().max() / want.abs().max()
        assert err <= 1e-5, (key, err)


def _vqvae_worker(rank, world, port, out_dir):
    from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        results = {}
        for mode in ("packed", "per_quantiser"):
            torch.manual_seed(0)                      # identical weights and codebooks on every rank
            model = VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2},
                          adapt_quantized_durations=False).to(dev).train()
            if mode == "per_quantiser":
                model.ema_exchange = None             # each quantiser all-reduces inside its forward
            launches = []
            if model.ema_exchange is not None:
                plain = model.ema_exchange.launch
                model.ema_exchange.launch = lambda: (launches.append(1), plain())[1]
            g = torch.Generator().manual_seed(100 + rank)
            for step in range(2):
                spec = torch.randn(2, 2, 256, 32, generator=g).to(dev)
                out = model(spec)
                out[0].mean().backward()
            results[mode] = {k: v.detach().cpu() for k, v in model.state_dict().items() if "quantize_" in k and "conv" not in k}
            results[mode + "_launches"] = len(launches)
        torch.save(results, os.path.join(out_dir, f"v{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_one_packed_allreduce_per_step_equals_per_quantiser_allreduces(tmp_path):
    """``VQVAE.forward`` in data-parallel training issues ONE all-reduce of the
    packed top+bottom statistics (overlapped with the decoder); codebooks stay bit-identical
    across ranks and equal what two per-quantiser all-reduces give."""
    world = 2
    mp.spawn(_vqvae_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"v{r}.pt") for r in range(world)]
    assert got[0]["packed_launches"] == 2 and got[1]["packed_launches"] == 2      # one per step
    for key in got[0]["packed"]:
        assert torch.equal(got[0]["packed"][key], got[1]["packed"][key]), key       # ranks identical
        torch.testing.assert_close(got[0]["packed"][key], got[0]["per_quantiser"][key], rtol=1e-6, atol=1e-7)
    assert float(got[0]["packed"]["quantize_b.cluster_size"].sum()) > 0            # the update happened
