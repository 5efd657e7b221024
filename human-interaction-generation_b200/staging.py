"""Asynchronous host -> device copies of the small per-iteration tensors (timesteps, caption ids, index maps).

`tensor.to(device)` from pageable host memory is cudaMemcpyAsync + cudaStreamSynchronize inside torch: the host blocks until
every kernel already queued on the stream has finished.  One such copy per training iteration is enough to keep the host
from running ahead of the GPU, and the GPU then idles while the host prepares and issues the next forward (0.5 ms of a
16 ms iteration at the C4 batch).  `stage()` copies through a small ring of pinned buffers instead; a slot is reused only
after the copy that last read it has completed (an event per slot)."""
import os
import threading

import torch

_SLOTS = 64
_rings = {}
_lock = threading.Lock()


_SLOT_BYTES = 4096


class _Ring:
    def __init__(self):
        # ONE pinned allocation for all slots: cudaHostAlloc costs milliseconds, and a slot-by-slot ring kept allocating
        # for its first _SLOTS uses — 13 iterations into a timed loop (seen as 24 ms iterations right after warm-up)
        arena = torch.empty(_SLOTS * _SLOT_BYTES, dtype=torch.uint8).pin_memory()
        self.bufs = [arena[k * _SLOT_BYTES:(k + 1) * _SLOT_BYTES] for k in range(_SLOTS)]
        self.events = [None] * _SLOTS
        self.i = 0


def stage(t, device, dtype=None):
    """Device copy of the CPU tensor `t` (optionally cast), without a host synchronisation.  Tensors already on `device`
    (or a CPU target) take the ordinary path."""
    device = torch.device(device)
    t = torch.as_tensor(t)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if device.type != "cuda" or t.device.type != "cpu" or t.numel() == 0 or os.environ.get("HIG_STAGE", "1") == "0":
        return t.to(device)        # (HIG_STAGE=0: the blocking copy, for A/B measurements)
    if t.is_pinned():
        return t.to(device, non_blocking=True)
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    with _lock:
        key = device.index if device.index is not None else torch.cuda.current_device()
        ring = _rings.get(key)
        if ring is None:
            ring = _rings[key] = _Ring()
        k = ring.i % _SLOTS
        ring.i += 1
        if ring.events[k] is not None:
            ring.events[k].synchronize()        # long done unless the host is a whole ring of copies ahead
        buf = ring.bufs[k]
        if buf.numel() < nbytes:             # a larger tensor than the slots were cut for: this slot gets its own buffer
            buf = ring.bufs[k] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        view = buf[:nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        out = view.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        ring.events[k] = ev
    return out
