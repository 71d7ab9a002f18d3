"""Sharded code extraction: the loop of the reference's ``extract_code.py:42-79`` on the
B200 path, calling ``encode`` only (the reference runs the decoder too and discards it,
extract_code.py:67) and without a host sync per note.

    audio batches --front end kernel--> spectrograms --VQVAE.encode_codes--> (top, bottom)

Rows keep the reference's ``CodeRow(top, bottom, attributes, filename)`` shape
(utils/datasets/lmdb_dataset.py:15); writing them to LMDB is left to the caller because
``lmdb`` is not available in this image -- ``save_shard`` writes an ``.npz`` per rank instead.
"""
from collections import namedtuple
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .utils import distributed as dist_utils

CodeRow = namedtuple('CodeRow', ['top', 'bottom', 'attributes', 'filename'])


class SpectrogramBatches:
    """On-GPU wav -> spectrogram batching with the call shape of the reference's
    ``WavToSpectrogramDataLoader`` (extract_code.py:199-206): iterating yields
    ``(spectrogram_batch, names)`` with the spectrograms computed on the helper's device.

    ``source`` yields ``(audio [b, T] float tensor on any device, names)``."""

    def __init__(self, source: Iterable[Tuple[torch.Tensor, Sequence[str]]], spectrograms_helper,
                 device: torch.device, transform: Optional[Callable] = None, prefetch: bool = True):
        self.source, self.helper, self.device, self.transform = source, spectrograms_helper, device, transform
        self.prefetch = prefetch
        self._side = None        # one copy stream per loader, so its allocator pool is reused
        # the helper may write 2x2 space-to-depth blocks (see SpectrogramsHelper); the encoder
        # has to be told (extract_codes reads this attribute)
        self.space_to_depth = bool(getattr(spectrograms_helper, "space_to_depth", False))

    def _upload(self, item, stream):
        audio, names = item
        with torch.cuda.stream(stream):
            dev_audio = audio.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(stream)
        return dev_audio, names, ready

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, Sequence[str]]]:
        """With ``prefetch`` the host->device copy of batch i+1 (pinned source memory) runs
        on a side stream while batch i is transformed and encoded."""
        main = torch.cuda.current_stream(self.device)
        if self.prefetch and self._side is None:
            self._side = torch.cuda.Stream(self.device)
        side = self._side if self.prefetch else main
        it = iter(self.source)
        pending = None
        for item in it:
            pending = self._upload(item, side)
            break
        while pending is not None:
            dev_audio, names, ready = pending
            pending = None
            for item in it:
                pending = self._upload(item, side)
                break
            main.wait_event(ready)
            dev_audio.record_stream(main)
            spec = self.helper.to_spectrogram(dev_audio)
            if self.transform is not None:
                spec = self.transform(spec)
            yield spec, names


def synthetic_source(n_notes: int, batch: int, rank: int = 0, world_size: int = 1,
                     n_samples: int = 64000, pin: bool = True):
    """This rank's contiguous shard of ``n_notes`` synthetic notes, in batches."""
    from .utils import synthetic
    start, stop = dist_utils.shard_range(n_notes, rank, world_size)
    for lo in range(start, stop, batch):
        hi = min(stop, lo + batch)
        audio = synthetic.synthetic_notes(hi - lo, n_samples=n_samples, seed=synthetic.AUDIO_SEED + lo)
        yield (audio.pin_memory() if pin else audio), [f"note_{i:07d}" for i in range(lo, hi)]


@torch.no_grad()
def extract_codes(loader: Iterable[Tuple[torch.Tensor, Sequence[str]]], model,
                  sink: Optional[Callable[[List[CodeRow]], None]] = None) -> List[CodeRow]:
    """extract_code.py:62-79 without the decode and with one device->host copy per batch.

    ``model`` exposes ``encode_codes(spec) -> (id_t, id_b)`` (this repo's VQVAE) or the
    reference's ``encode`` 7-tuple.  Returns (and optionally streams to ``sink``) the rows."""
    rows: List[CodeRow] = []
    pending = None
    # Two pinned staging pairs, reused: allocating page-locked memory per batch is slow and is
    # an implicit device synchronisation (it breaks the copy/compute overlap).  flush() copies
    # the maps out of the staging pair before that pair's next use, two batches later.
    staging: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None, None]

    def staging_pair(slot, id_t, id_b):
        pair = staging[slot]
        if pair is None or pair[0].numel() < id_t.numel() or pair[1].numel() < id_b.numel():
            pair = (torch.empty(id_t.numel(), dtype=id_t.dtype, pin_memory=True),
                    torch.empty(id_b.numel(), dtype=id_b.dtype, pin_memory=True))
            staging[slot] = pair
        return (pair[0][:id_t.numel()].view(id_t.shape), pair[1][:id_b.numel()].view(id_b.shape))

    def flush(item):
        host_t, host_b, names, done = item
        done.synchronize()
        tops, bottoms = host_t.numpy().copy(), host_b.numpy().copy()
        batch_rows = [CodeRow(top=t, bottom=b, attributes={}, filename=n)
                      for t, b, n in zip(tops, bottoms, names)]
        if sink is not None:
            sink(batch_rows)
        rows.extend(batch_rows)

    model.eval()
    s2d = bool(getattr(loader, "space_to_depth", False))
    for step, (spec, names) in enumerate(loader):
        if hasattr(model, "encode_codes"):
            id_t, id_b = model.encode_codes(spec, space_to_depth=True) if s2d else model.encode_codes(spec)
        else:
            if s2d:
                raise ValueError("space-to-depth spectrograms need this repo's VQVAE.encode_codes")
            out = model.encode(spec)
            id_t, id_b = out[3], out[4]
        host_t, host_b = staging_pair(step % 2, id_t, id_b)
        host_t.copy_(id_t, non_blocking=True)
        host_b.copy_(id_b, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        if pending is not None:          # the previous batch's copy overlaps this batch's compute
            flush(pending)
        pending = (host_t, host_b, list(names), done)
    if pending is not None:
        flush(pending)
    return rows


def save_shard(rows: Sequence[CodeRow], path) -> None:
    """One ``.npz`` per rank: names, top and bottom code maps (idempotent per note name,
    like the reference's ``dupsort=False`` LMDB keys, extract_code.py:47-50)."""
    np.savez_compressed(path, filename=np.array([r.filename for r in rows]),
                        top=np.stack([r.top for r in rows]) if rows else np.zeros((0,)),
                        bottom=np.stack([r.bottom for r in rows]) if rows else np.zeros((0,)))
