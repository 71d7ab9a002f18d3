"""World-size-2 CPU (gloo) checks of the N>1 host logic:
 * the packed EMA statistics all-reduced over ranks give the update one process computes on
   the concatenated batch (the parity target of SURVEY.md F3);
 * note shards partition the work with no duplicate and no gap."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interactive_spectrogram_inpainting_b200.utils import distributed as du
from interactive_spectrogram_inpainting_b200.utils import synthetic
from interactive_spectrogram_inpainting_b200.vqvae.bottleneck import QuantizedBottleneck
from oracle import quantizer_oracle as qo

WORLD = 2
DIM, K, ROWS = 16, 40, 300


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _packed_stats(rows, ind):
    """[counts | embed_sum code-major], the layout isi_vq_gather_stats produces."""
    onehot = torch.nn.functional.one_hot(ind, K).float()
    return torch.cat([onehot.sum(0), (onehot.t() @ rows).reshape(-1)])


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        embed = synthetic.synthetic_codebook(DIM, K)
        full = synthetic.synthetic_features(ROWS, embed, 77)
        lo, hi = du.shard_range(ROWS, rank, WORLD)
        assert (rank, WORLD) == du.world() and du.is_master_process() == (rank == 0)
        mine = full[lo:hi]
        ind = qo.assign(mine, embed)
        module = QuantizedBottleneck(DIM, K)          # host logic only: never touches CUDA here
        stats = module.reduce_ema_stats(_packed_stats(mine, ind))
        counts, embed_sum = stats[:K], stats[K:].view(K, DIM).t().contiguous()
        st = qo.CodebookState(embed.clone(), torch.zeros(K), embed.clone())
        qo.ema_update(st, mine, ind, 0.99, 1e-5, counts=counts, embed_sum=embed_sum)
        torch.save({"embed": st.embed, "cluster_size": st.cluster_size, "embed_avg": st.embed_avg,
                    "shard": (lo, hi)}, os.path.join(out_dir, f"rank{rank}.pt"))
        module.sync_ema_stats = False                  # opt-out keeps statistics local
        local = _packed_stats(mine, ind)
        assert torch.equal(module.reduce_ema_stats(local.clone()), local)
    finally:
        dist.destroy_process_group()


def test_allreduced_ema_equals_single_process_on_concatenated_batch(tmp_path):
    mp.spawn(_worker, args=(_free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(WORLD)]
    embed = synthetic.synthetic_codebook(DIM, K)
    full = synthetic.synthetic_features(ROWS, embed, 77)
    st = qo.CodebookState(embed.clone(), torch.zeros(K), embed.clone())
    qo.ema_update(st, full, qo.assign(full, embed), 0.99, 1e-5)
    for g in got:
        torch.testing.assert_close(g["cluster_size"], st.cluster_size, rtol=1e-6, atol=0)
        torch.testing.assert_close(g["embed_avg"], st.embed_avg, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(g["embed"], st.embed, rtol=1e-5, atol=1e-6)
    # every rank ends with the identical codebook, and the shards tile the batch
    assert torch.equal(got[0]["embed"], got[1]["embed"])
    assert got[0]["shard"][1] == got[1]["shard"][0] and got[1]["shard"][1] == ROWS
