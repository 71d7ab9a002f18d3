"""BASELINE config 3: one VQ-VAE-2 training step (train_vqvae.py:169-192) per rank, batch 64 per
GPU, EMA codebook statistics all-reduced over NCCL (this repo's quantiser inside the torch conv
encoder/decoder, DDP for the conv gradients, Adam).

    python tools/bench_train.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_train.py                  # 8 GPUs

Prints one JSON line: ms per step (CUDA events, max over ranks), notes/s over all ranks, and
the share of the step spent in this repo's C-ABI calls and in the EMA all-reduce (timed with
CUDA events around the collective; the convolutions are torch/cuDNN and not this repo's code).
"""
import json
import os
import pathlib
import statistics
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from interactive_spectrogram_inpainting_b200 import _lib  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils.spectrograms_helper import MelSpectrogramsHelper  # noqa: E402
from interactive_spectrogram_inpainting_b200.utils import synthetic  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae import bottleneck  # noqa: E402
from interactive_spectrogram_inpainting_b200.vqvae.vqvae import VQVAE  # noqa: E402

BATCH, STEPS, WARMUP = 64, 10, 3


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    helper = MelSpectrogramsHelper(channels_last=True).to(dev)
    model = VQVAE(in_channel=2, resolution_factors={"bottom": 16, "top": 2},
                  adapt_quantized_durations=False).to(dev).to(memory_format=torch.channels_last).train()
    net = model
    if world > 1:       # the quantiser keeps its own buffers in sync: no per-forward broadcast
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False)
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    audio = synthetic.synthetic_notes(BATCH, seed=synthetic.AUDIO_SEED + rank).to(dev)

    # time the EMA all-reduce separately: wrap the module's collective
    reduce_events = []
    plain_reduce = bottleneck.QuantizedBottleneck.reduce_ema_stats

    def timed_reduce(self, stats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = plain_reduce(self, stats)
        b.record()
        reduce_events.append((a, b))
        return out
    bottleneck.QuantizedBottleneck.reduce_ema_stats = timed_reduce

    def step():
        spec = helper.to_spectrogram(audio)                     # front end in the loop, like train()
        recon, diff, *_ = net(spec)
        loss = F.mse_loss(recon, spec) + 0.25 * diff.mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(WARMUP):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reduce_events.clear()
    _lib.event_log = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(STEPS):
        loss = step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    calls, _lib.event_log = _lib.event_log, None
    ms = torch.tensor([t0.elapsed_time(t1) / STEPS], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per_call = {}
    for name, a, b in calls:
        per_call[name] = per_call.get(name, 0.0) + a.elapsed_time(b) / STEPS
    reduce_ms = sum(a.elapsed_time(b) for a, b in reduce_events) / STEPS
    # codebooks must stay identical across ranks (the point of the all-reduce)
    same = True
    if world > 1:
        ref = model.quantize_b.embed.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([float(torch.equal(ref, model.quantize_b.embed))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    if rank == 0:
        print(json.dumps({
            "config": "cfg3: VQ-VAE-2 training step, batch 64 per GPU, EMA statistics all-reduced (NCCL)",
            "n_gpus": world, "ms_per_step": ms.item(), "notes_per_s": world * BATCH / (ms.item() * 1e-3),
            "loss": float(loss), "isi_calls_ms_per_step": {k: round(v, 4) for k, v in per_call.items()},
            "isi_total_ms_per_step": round(sum(per_call.values()), 4),
            "ema_allreduce_ms_per_step": round(reduce_ms, 4),
            "ema_allreduce_bytes": 2 * 4 * 512 * 65 if world > 1 else 0,
            "codebooks_identical_across_ranks": same}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
