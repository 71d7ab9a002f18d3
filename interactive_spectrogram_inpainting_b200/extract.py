"""Sharded code extraction: the loop of the reference's ``extract_code.py:42-79`` on the
B200 path, calling ``encode`` only (the reference runs the decoder too and discards it,
extract_code.py:67) and without a host sync per note.

    audio batches --front end kernel--> spectrograms --VQVAE.encode_codes--> (top, bottom)

Rows keep the reference's ``CodeRow(top, bottom, attributes, filename)`` shape
(utils/datasets/lmdb_dataset.py:15).  ``row_record`` serialises a row to exactly the key / value
bytes the reference stores (extract_code.py:71-79), ``write_lmdb`` commits a batch of them in ONE
transaction (the reference opens one per note) when the ``lmdb`` package is installed -- it is not
in this image, where ``save_shard`` writes an ``.npz`` per rank instead.
"""
from collections import namedtuple
from collections.abc import Mapping
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import pickle
import threading

import numpy as np
import torch

from .utils import distributed as dist_utils

CodeRow = namedtuple('CodeRow', ['top', 'bottom', 'attributes', 'filename'])


def _split_item(item):
    """A source item is ``(audio, names)`` or ``(audio, names, attributes)`` with ``attributes``
    a mapping {attribute name: per-note values} -- the label-encoded categorical fields the
    reference's loader yields next to each batch (extract_code.py:62-63,
    ``--categorical_fields`` :104-105)."""
    if len(item) == 2:
        return item[0], item[1], None
    audio, names, attributes = item
    return audio, names, (attributes if attributes else None)


def _row_attributes(attributes, index: int) -> dict:
    """``dict(zip(attribute_names, attributes))`` of extract_code.py:73-74 for one note: 0-dim
    int64 tensors, because the reference's reader calls ``.view(1)`` on them
    (utils/datasets/lmdb_dataset.py:84-86)."""
    if attributes is None:
        return {}
    return {name: torch.as_tensor(values[index]).detach().cpu().reshape(()).clone()
            for name, values in attributes.items()}


def _unpack_batch(batch):
    """A loader batch is ``(spec, names)``, ``(spec, names, attributes)`` or the reference
    loader's ``(spec, *categorical, attributes_batch)`` with ``attributes_batch['note_str']``
    the names and ``attributes_batch['categorical_fields']`` (ours) the field names."""
    if isinstance(batch[-1], Mapping) and 'note_str' in batch[-1]:
        info = batch[-1]
        fields = list(info.get('categorical_fields', range(len(batch) - 2)))
        return batch[0], info['note_str'], dict(zip(fields, batch[1:-1])) or None
    return _split_item(batch)


class SpectrogramBatches:
    """On-GPU wav -> spectrogram batching with the call shape of the reference's
    ``WavToSpectrogramDataLoader`` (extract_code.py:199-206): iterating yields
    ``(spectrogram_batch, names)`` with the spectrograms computed on the helper's device.

    ``source`` yields ``(audio [b, T] float tensor on any device, names)`` or ``(audio, names,
    attributes)`` (see ``_split_item``).  With ``reference_protocol`` a batch comes out as the
    reference's loader yields it (extract_code.py:62-63): ``(spectrogram_batch,
    *categorical_attribute_tensors, attributes_batch)`` with ``attributes_batch['note_str']`` the
    names; otherwise ``(spectrogram_batch, names[, attributes])``."""

    def __init__(self, source: Iterable[Tuple[torch.Tensor, Sequence[str]]], spectrograms_helper,
                 device: torch.device, transform: Optional[Callable] = None, prefetch: bool = True,
                 reference_protocol: bool = False):
        self.source, self.helper, self.device, self.transform = source, spectrograms_helper, device, transform
        self.prefetch = prefetch
        self.reference_protocol = reference_protocol
        self._side = None        # one copy stream per loader, so its allocator pool is reused
        # the helper may write 2x2 space-to-depth blocks (see SpectrogramsHelper); the encoder
        # has to be told (extract_codes reads this attribute)
        self.space_to_depth = getattr(spectrograms_helper, "space_to_depth", False)     # False / True / "transposed"

    def _upload(self, item, stream):
        audio, names, attributes = _split_item(item)
        with torch.cuda.stream(stream):
            dev_audio = audio.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(stream)
        return dev_audio, names, attributes, ready

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
            dev_audio, names, attributes, ready = pending
            pending = None
            for item in it:
                pending = self._upload(item, side)
                break
            main.wait_event(ready)
            dev_audio.record_stream(main)
            spec = self.helper.to_spectrogram(dev_audio)
            if self.transform is not None:
                spec = self.transform(spec)
            if self.reference_protocol:
                fields = list(attributes) if attributes else []
                yield (spec, *[torch.as_tensor(attributes[f]) for f in fields],
                       {'note_str': list(names), 'categorical_fields': fields})
            elif attributes is None:
                yield spec, names
            else:
                yield spec, names, attributes


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
        host_t, host_b, names, attributes, done = item
        done.synchronize()
        tops, bottoms = host_t.numpy().copy(), host_b.numpy().copy()
        batch_rows = [CodeRow(top=t, bottom=b, attributes=_row_attributes(attributes, i), filename=n)
                      for i, (t, b, n) in enumerate(zip(tops, bottoms, names))]
        if sink is not None:
            sink(batch_rows)
        rows.extend(batch_rows)

    model.eval()
    s2d = getattr(loader, "space_to_depth", False)            # False / True / "transposed"
    for step, batch in enumerate(loader):
        spec, names, attributes = _unpack_batch(batch)
        if hasattr(model, "encode_codes"):
            id_t, id_b = model.encode_codes(spec, space_to_depth=s2d) if s2d else model.encode_codes(spec)
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
        pending = (host_t, host_b, list(names), attributes, done)
    if pending is not None:
        flush(pending)
    return rows


class _GraphedStep:
    """Front end + ``encode_codes`` for one batch shape, captured in a CUDA graph: a step is ~30
    kernel launches through Python and ctypes; replaying them costs the host one call, which is
    what matters when eight ranks share a host's cores.  The graph holds the addresses of the
    prepared codebooks and projection weights: eval mode, weights that no longer change."""

    def __init__(self, helper, model, example: torch.Tensor, space_to_depth, warmup: int = 3):
        device = example.device
        self.audio = torch.empty_like(example)
        self.audio.copy_(example)

        def run():
            spec = helper.to_spectrogram(self.audio)
            return model.encode_codes(spec, space_to_depth=space_to_depth) if space_to_depth else model.encode_codes(spec)
        stream = torch.cuda.Stream(device)
        stream.wait_stream(torch.cuda.current_stream(device))
        with torch.no_grad(), torch.cuda.stream(stream):
            for _ in range(max(1, warmup)):      # cuDNN plan selection and cache fills stay outside
                run()
        torch.cuda.current_stream(device).wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (NCCL watchdog, loader workers) may call the
        # CUDA runtime while this thread captures
        with torch.no_grad(), torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.id_t, self.id_b = run()

    def __call__(self, audio: torch.Tensor):
        self.audio.copy_(audio, non_blocking=True)
        self.graph.replay()
        return self.id_t, self.id_b


class CodeExtractor:
    """``extract_code.py:62-79`` from audio batches to ``CodeRow``s as one pipeline:

        pinned host audio --H2D (side stream, one batch ahead)--> front end + encode_codes
        (one CUDA-graph replay per batch when ``cuda_graph``) --D2H--> rows

    ``run(source)`` takes what ``SpectrogramBatches`` takes: an iterable of ``(audio [b, T],
    names)`` or ``(audio, names, attributes)``.  Graphs are captured per batch shape on first use and kept for later calls; a
    shape whose capture fails (or ``cuda_graph=False``) runs the same calls eagerly.  The model
    must be this repo's ``VQVAE`` in eval mode with fixed weights."""

    def __init__(self, spectrograms_helper, model, device: torch.device, cuda_graph: bool = True,
                 prefetch_depth: int = 2):
        if not hasattr(model, "encode_codes"):
            raise TypeError("CodeExtractor needs this repo's VQVAE (encode_codes)")
        self.helper, self.model, self.device = spectrograms_helper, model.eval(), device
        self.cuda_graph = cuda_graph
        self.space_to_depth = getattr(spectrograms_helper, "space_to_depth", False)     # False / True / "transposed"
        self._graphs = {}
        self._side = torch.cuda.Stream(device)
        # uploads run `prefetch_depth` batches ahead of the compute (two: a slow copy -- eight
        # ranks share one host's memory and PCIe switches -- does not stall the next step)
        self.prefetch_depth = max(1, int(prefetch_depth))
        self._dev_audio = [None] * (self.prefetch_depth + 1)
        # code maps leave the device as int32 (half the bytes of the int64 the kernels write;
        # widened again on the host) on their own stream: the main stream goes straight on to
        # the next batch instead of waiting ~0.1 ms for 2.3 MB to cross PCIe
        self._side_out = torch.cuda.Stream(device)
        self._dev_codes = [None, None]
        self._host_codes = [None, None]
        self.d2h_bytes_last_batch = 0
        self.graph_failures: List[str] = []

    def _step(self, audio: torch.Tensor):
        key = (tuple(audio.shape), audio.dtype)
        graphed = self._graphs.get(key)
        if graphed is None and self.cuda_graph:
            try:
                graphed = _GraphedStep(self.helper, self.model, audio, self.space_to_depth)
            except Exception as exc:       # capture is an optimisation: the eager calls are the same work
                self.graph_failures.append(f"{key}: {exc}")
                graphed = False
            self._graphs[key] = graphed
        if graphed:
            return graphed(audio)
        spec = self.helper.to_spectrogram(audio)
        return (self.model.encode_codes(spec, space_to_depth=self.space_to_depth) if self.space_to_depth
                else self.model.encode_codes(spec))

    @torch.no_grad()
    def run(self, source: Iterable[Tuple[torch.Tensor, Sequence[str]]],
            sink: Optional[Callable[[List[CodeRow]], None]] = None) -> List[CodeRow]:
        main, side = torch.cuda.current_stream(self.device), self._side
        n_slots = self.prefetch_depth + 1
        consumed = [None] * n_slots      # per slot: the step that read the device copy is enqueued
        rows: List[CodeRow] = []

        def upload(item, slot):
            audio, names, attributes = _split_item(item)
            buf = self._dev_audio[slot]
            with torch.cuda.stream(side):
                if buf is None or buf.shape != audio.shape or buf.dtype != audio.dtype:
                    # a new block may be memory the main stream's previous step has just freed
                    # and is still reading: the copy below must not overtake that step
                    side.wait_stream(main)
                    buf = torch.empty(audio.shape, dtype=audio.dtype, device=self.device)
                    buf.record_stream(main)
                    self._dev_audio[slot] = buf
                if consumed[slot] is not None:
                    side.wait_event(consumed[slot])
                buf.copy_(audio, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(side)
            return buf, list(names), attributes, ready

        def code_buffers(slot, id_t, id_b):
            pair = self._host_codes[slot]
            if pair is None or pair[0].shape != id_t.shape or pair[1].shape != id_b.shape:
                pair = (torch.empty(id_t.shape, dtype=torch.int32, pin_memory=True),
                        torch.empty(id_b.shape, dtype=torch.int32, pin_memory=True))
                self._host_codes[slot] = pair
                self._dev_codes[slot] = (torch.empty(id_t.shape, dtype=torch.int32, device=self.device),
                                         torch.empty(id_b.shape, dtype=torch.int32, device=self.device))
            return self._dev_codes[slot], pair

        def flush(item):
            host_t, host_b, names, attributes, done = item
            done.synchronize()
            tops, bottoms = host_t.numpy().astype(np.int64), host_b.numpy().astype(np.int64)
            batch_rows = [CodeRow(top=t, bottom=b, attributes=_row_attributes(attributes, i), filename=n)
                          for i, (t, b, n) in enumerate(zip(tops, bottoms, names))]
            if sink is not None:
                sink(batch_rows)
            rows.extend(batch_rows)

        from collections import deque
        it = iter(source)
        uploads, next_slot = deque(), 0

        def top_up():
            # a new upload takes the slot of the step BEFORE the current one: already enqueued,
            # its `consumed` event recorded
            nonlocal next_slot
            while len(uploads) < self.prefetch_depth:
                item = next(it, None)
                if item is None:
                    return
                uploads.append(upload(item, next_slot) + (next_slot,))
                next_slot = (next_slot + 1) % n_slots

        top_up()
        pending, step = None, 0
        while uploads:
            audio, names, attributes, ready, slot = uploads.popleft()
            top_up()
            main.wait_event(ready)
            id_t, id_b = self._step(audio)
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(main)
            # (a slot's buffers are free again: the batch that used them two steps ago was
            # flushed -- host-synchronised -- during the previous iteration)
            (dev_t, dev_b), (host_t, host_b) = code_buffers(step % 2, id_t, id_b)
            dev_t.copy_(id_t)            # int64 -> int32 on the device, before the next replay
            dev_b.copy_(id_b)            # overwrites the graph's output buffers
            narrowed = torch.cuda.Event()
            narrowed.record(main)
            with torch.cuda.stream(self._side_out):
                self._side_out.wait_event(narrowed)
                host_t.copy_(dev_t, non_blocking=True)
                host_b.copy_(dev_b, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._side_out)
            self.d2h_bytes_last_batch = 4 * (host_t.numel() + host_b.numel())
            if pending is not None:      # the previous batch's copy overlaps this batch's compute
                flush(pending)
            pending = (host_t, host_b, names, attributes, done)
            step += 1
        if pending is not None:
            flush(pending)
        return rows


# The reference pickles ``CodeRow`` instances of the class defined in its own module; its readers
# (utils/datasets/lmdb_dataset.py:79-89, used by train_autoregressive_model.py:482-485 and
# flask_server.py:273-279) unpickle through that import path.  Records written here name the
# same path, so they load there without this package being importable.
_REFERENCE_ROW_MODULE = "interactive_spectrogram_inpainting.utils.datasets.lmdb_dataset"
_ReferenceCodeRow = namedtuple('CodeRow', ['top', 'bottom', 'attributes', 'filename'])
_ReferenceCodeRow.__module__ = _REFERENCE_ROW_MODULE
_ReferenceCodeRow.__qualname__ = 'CodeRow'


_record_lock = threading.Lock()     # row_record lends sys.modules a placeholder while it pickles


def row_record(row: CodeRow) -> Tuple[bytes, bytes]:
    """``(key, value)`` as ``extract_code.py:71-79`` stores a note: key = the note name in UTF-8,
    value = ``pickle.dumps`` of the reference's ``CodeRow`` (``top`` / ``bottom`` int64 arrays,
    ``attributes`` dict, ``filename``)."""
    import sys
    import types
    record = _ReferenceCodeRow(top=np.asarray(row.top), bottom=np.asarray(row.bottom),
                               attributes=dict(row.attributes), filename=row.filename)
    # pickle resolves the class by importing its module; lend it a placeholder module holding
    # the class when the reference package is not importable here
    with _record_lock:
        lent = []
        loaded = sys.modules.get(_REFERENCE_ROW_MODULE)
        if loaded is None:
            # pickle imports the module and its parent packages to verify the name: lend them
            parts = _REFERENCE_ROW_MODULE.split(".")
            for i in range(1, len(parts) + 1):
                name = ".".join(parts[:i])
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    lent.append(name)
            sys.modules[_REFERENCE_ROW_MODULE].CodeRow = _ReferenceCodeRow
        elif getattr(loaded, "CodeRow", None) is not _ReferenceCodeRow:
            record = loaded.CodeRow(*record)          # the real reference module: pickle its own class
        try:
            value = pickle.dumps(record)
        finally:
            for name in lent:
                del sys.modules[name]
    return row.filename.encode('utf-8'), value


def write_lmdb(rows: Sequence[CodeRow], env, db=None) -> int:
    """Commit ``rows`` to an open ``lmdb.Environment`` in ONE write transaction (the reference
    opens one per note, extract_code.py:76-78); ``db`` as returned by ``env.open_db(b'codes',
    dupsort=False)``.  Existing keys are overwritten, like the reference's ``put``.  Returns the
    number of records written.  Needs the ``lmdb`` package only through ``env``."""
    records = [row_record(r) for r in rows]
    with env.begin(db=db, write=True) as txn:
        for key, value in records:
            txn.put(key, value)
    return len(records)


def write_label_encoders(env, label_encoders) -> None:
    """``extract_code.py:52-57``: the label encoders ride along in the database's unnamed
    default table under the key ``label_encoders`` (master process only, like the reference)."""
    if not dist_utils.is_master_process():
        return
    with env.begin(write=True) as txn:
        txn.put('label_encoders'.encode('utf-8'), pickle.dumps(label_encoders))


def save_shard(rows: Sequence[CodeRow], path) -> None:
    """One ``.npz`` per rank: names, top and bottom code maps (idempotent per note name,
    like the reference's ``dupsort=False`` LMDB keys, extract_code.py:47-50)."""
    np.savez_compressed(path, filename=np.array([r.filename for r in rows]),
                        top=np.stack([r.top for r in rows]) if rows else np.zeros((0,)),
                        bottom=np.stack([r.bottom for r in rows]) if rows else np.zeros((0,)))
