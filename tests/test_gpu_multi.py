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
