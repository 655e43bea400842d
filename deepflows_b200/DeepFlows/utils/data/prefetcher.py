"""Device-side input pipeline (new). The reference's loop hands every numpy batch to
`Tensor(batch, device=cuda)`, i.e. a blocking `cudaMemcpy` from pageable memory right before the step
(DeepFlows/utils/data/dataloader.py:60-139 + ndarray_backend_cuda.cu:700-716). `DevicePrefetcher` wraps any
iterable of `(x, y)` numpy batches (a `DataLoader`) and yields device-resident `Tensor`s instead, with the
copy of batch i+1 running on the copy stream while the caller computes on batch i:

    for x, t in DevicePrefetcher(loader):                 # Tensors on the cuda device
        loss = criterion(model(x), t); ...

    x, t = Tensor(...), Tensor(...)                       # static input buffers of a CapturedStep
    for _ in DevicePrefetcher(loader, into=(x, t)):       # each batch is moved into x / t
        loss = step()

Batches go host batch -> pinned buffer (one memcpy) -> device staging buffer (`dfb_prefetch_from_host`, copy
stream) -> consumer. Two pinned / staging buffer pairs alternate, so the host can fill the next pinned buffer
while the previous copy is still in flight. Ragged last batches are supported (buffers are sized for the
largest batch seen; `into` requires equal shapes).
"""
import numpy as np

from ... import backend_api
from ...backend.backend_tensor import BackendTensor
from ...tensor import Tensor


class DevicePrefetcher:
    def __init__(self, loader, device=None, into=None):
        self.loader = loader
        self.device = device if device is not None else backend_api.cuda()
        self.into = into
        self._slots = [None, None]  # per slot: [(pinned ndarray, device Array, capacity)] per field
        self._consumed = [None, None]  # per slot: event recorded on the compute stream after the slot was delivered

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot, fields):
        """Copy the numpy fields into the slot's pinned buffers and enqueue their host->device prefetch."""
        dev = self.device
        if self._consumed[slot] is not None:
            # the host is about to overwrite this slot's pinned buffers: its previous copy must have left them
            # (the event sits behind that copy's consumer on the compute stream)
            dev.event_synchronize(self._consumed[slot])
        bufs = self._slots[slot]
        if bufs is None or any(b[2] < f.size for b, f in zip(bufs, fields)):
            bufs = [(dev.pinned_empty(f.size), dev.Array(f.size), f.size) for f in fields]
            self._slots[slot] = bufs
        for (pinned, arr, _), f in zip(bufs, fields):
            pinned[:f.size] = f.reshape(-1)
            dev.prefetch_from_pinned(pinned, arr, f.size)
        return [f.shape for f in fields]

    def _deliver(self, slot, shapes):
        out = self._deliver_inner(slot, shapes)
        if self._consumed[slot] is None:
            self._consumed[slot] = self.device.event_create()
        self.device.event_record(self._consumed[slot])
        return out

    def __del__(self):
        for ev in self._consumed:
            if ev is not None:
                try:
                    self.device.event_destroy(ev)
                except Exception:
                    pass

    def _deliver_inner(self, slot, shapes):
        dev = self.device
        dev.prefetch_wait()  # the compute stream now sees the staged batch
        bufs = self._slots[slot]
        if self.into is not None:
            for target, (_, arr, _), shape in zip(self.into, bufs, shapes):
                if tuple(target.shape) != tuple(shape):
                    raise ValueError("DevicePrefetcher(into=...): batch shape %s does not match the buffer %s"
                                     % (tuple(shape), tuple(target.shape)))
                data = target.data if target.data.is_compact() else None
                if data is None:
                    raise ValueError("DevicePrefetcher(into=...): target tensors must be compact")
                dev.copy(arr, (data._handle, data._offset), int(np.prod(shape)))
            return self.into
        out = []
        for (_, arr, _), shape in zip(bufs, shapes):
            n = int(np.prod(shape))
            fresh = dev.Array(n)          # the staging buffer is reused two batches later: hand out a copy
            dev.copy(arr, fresh, n)
            out.append(Tensor(BackendTensor.make(tuple(shape), device=dev, handle=fresh)))
        return tuple(out)

    def __iter__(self):
        pending = None  # (slot, shapes) of the batch whose copy is in flight
        slot = 0
        for batch in self.loader:
            fields = [np.ascontiguousarray(f, dtype=np.float32) for f in batch]
            if pending is not None:
                # deliver batch i (its copy was enqueued one iteration ago), then start batch i+1's copy: the
                # prefetch waits only for what is on the compute stream now, i.e. the device-to-device moves
                # of _deliver, not for the step the caller is about to run
                delivered = self._deliver(*pending)
                pending = (slot, self._stage(slot, fields))
                slot ^= 1
                yield delivered
            else:
                pending = (slot, self._stage(slot, fields))
                slot ^= 1
        if pending is not None:
            yield self._deliver(*pending)
