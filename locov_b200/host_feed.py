"""Host -> device staging of input batches for the drop-in heads.

The reference moves every batch to the device inside the training step (``.to(self.device)``; distill_prop_mmss_gcnn.py:273-399 hands
the heads CUDA tensors), so the 28 MB of RoI features of an LSM batch cross PCIe while the GPU waits.  ``HostFeed`` is the runtime
piece that removes that wait without changing what the heads see: batches (dicts / tuples of PINNED host tensors) are copied on a
dedicated copy stream into a small ring of device buffers, and the copy of batch i + 1 runs while the kernels of batch i execute.

    feed = HostFeed(device)
    feed.submit(first)                       # enqueue H2D of the first batch
    for nxt in rest:
        feed.submit(nxt)                     # H2D of the next batch ...
        inputs = feed.take()                 # ... overlaps everything launched on the current stream from here on
        out = head(*inputs)

Ordering is carried by CUDA events only (no host synchronisation): ``take`` makes the current stream wait for the batch's copy;
``submit`` makes the copy stream wait for everything launched so far before it overwrites the ring slot that was handed out ``depth``
batches ago.
"""
from collections import deque

import torch

from ._lib import LocoError


def _map(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map(v, fn) for v in obj)
    return obj


def _tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, dict):
        for k in obj:
            yield from _tensors(obj[k])
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors(v)


def _signature(obj):
    """Structure of a batch: container kinds / keys and (shape, dtype) of every tensor."""
    if torch.is_tensor(obj):
        return (tuple(obj.shape), obj.dtype)
    if isinstance(obj, dict):
        return tuple((k, _signature(v)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return tuple(_signature(v) for v in obj)
    return None


class HostFeed:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise LocoError("HostFeed stages batches for a CUDA device")
        if depth < 2:
            raise LocoError("HostFeed needs at least two ring slots to overlap a copy with compute")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = [None] * depth            # (signature, device-side mirror of the batch structure)
        self._ready = deque()                   # (slot index, copy-done event) in submission order
        self._n = 0
        self.bytes_copied = 0

    def submit(self, batch):
        if len(self._ready) >= self.depth:
            raise LocoError(f"HostFeed: {self.depth} batches already in flight; take() one first")
        for t in _tensors(batch):
            if t.is_cuda or not t.is_pinned():
                raise LocoError("HostFeed.submit expects PINNED host tensors (an unpinned copy is synchronous and cannot overlap)")
        i = self._n % self.depth
        self._n += 1
        free = torch.cuda.Event()
        free.record(torch.cuda.current_stream(self.device))      # everything that may still read this slot has been launched by now
        self.copy_stream.wait_event(free)
        sig = _signature(batch)
        if self._slots[i] is None or self._slots[i][0] != sig:
            def alloc(t):
                d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                d.record_stream(self.copy_stream)         # written on the copy stream, freed (if ever) on the allocating one
                return d
            self._slots[i] = (sig, _map(batch, alloc))
        with torch.cuda.stream(self.copy_stream):
            for d, s in zip(_tensors(self._slots[i][1]), _tensors(batch)):
                d.copy_(s, non_blocking=True)
                self.bytes_copied += s.numel() * s.element_size()
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._ready.append((i, done))

    def take(self):
        if not self._ready:
            raise LocoError("HostFeed.take: nothing submitted")
        i, done = self._ready.popleft()
        torch.cuda.current_stream(self.device).wait_event(done)
        return self._slots[i][1]
