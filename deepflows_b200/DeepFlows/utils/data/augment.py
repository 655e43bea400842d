"""Per-batch preparation on the device (new).

The reference's scripts prepare every batch in numpy on the host, between the loader and `Tensor(batch, device=cuda)`:
`augment_batch` (reference: test/ResNet_CIFAR10_cuda.py:129-148 - reflect padding, random crop, horizontal flip,
random erasing, clip) and one-hot targets with label smoothing (same file, lines 181-183). At a millisecond per
step those host loops are the bottleneck, so here the random DRAWS stay on the host - numpy's global generator in the
reference's order, a seeded run sees the reference's numbers - and travel with the batch as a small float table,
while the per-pixel work is one kernel over the device-resident batch (`dfb_augment_batch`, `dfb_onehot_smooth`).
Results are bit-identical to the numpy code.

    aug = BatchAugment(pad=4)
    loader = ((x, y, aug.draw(len(x), 32, 32, epoch, num_epochs)) for x, y in data_loader(...))
    for x, y, table in DevicePrefetcher(loader):             # all three on the device
        x = aug(x, table)
        t = smooth_one_hot(y, 10, eps=0.05)
        loss = criterion(model(x), t)
"""
import numpy as np

from ...backend.backend_tensor import BackendTensor
from ...tensor import Tensor

FIELDS = 8  # crop_y, crop_x, flip, erase_y, erase_x, erase_h, erase_w, unused (DFB_AUGMENT_FIELDS)


def _backend(t):
    return t.data if isinstance(t, Tensor) else t


class BatchAugment:
    """`draw` consumes numpy's global generator exactly like the reference's augment_batch; `__call__` applies a
    table of draws to a batch: a kernel for tensors on the cuda device, numpy for host arrays."""

    def __init__(self, pad=4, flip_p=0.5, erase_p=0.2, erase_frac=(0.1, 0.2), erase_off_epochs=5, clip=(-1.0, 1.0)):
        self.pad = int(pad)
        self.flip_p = float(flip_p)
        self.erase_p = float(erase_p)
        self.erase_frac = (float(erase_frac[0]), float(erase_frac[1]))
        self.erase_off_epochs = int(erase_off_epochs)
        self.clip = None if clip is None else (float(clip[0]), float(clip[1]))

    def draw(self, n, h, w, epoch=0, num_epochs=0):
        """The random numbers of one batch as an (n, 8) float32 table. Order of the draws [131-143]: crop rows,
        crop columns, flips, then - only while `epoch < num_epochs - erase_off_epochs` - one uniform that decides
        whether this batch is erased and, if so, the rectangle's height, width, rows and columns."""
        table = np.zeros((n, FIELDS), dtype=np.float32)
        span = 2 * self.pad + 1
        table[:, 0] = np.random.randint(0, span, size=n)
        table[:, 1] = np.random.randint(0, span, size=n)
        table[:, 2] = np.random.rand(n) < self.flip_p
        if epoch < num_epochs - self.erase_off_epochs and np.random.rand() < self.erase_p:
            lo, hi = self.erase_frac
            eh = max(1, int(h * np.random.uniform(lo, hi)))
            ew = max(1, int(w * np.random.uniform(lo, hi)))
            table[:, 3] = np.random.randint(0, h - eh + 1, size=n)
            table[:, 4] = np.random.randint(0, w - ew + 1, size=n)
            table[:, 5] = eh
            table[:, 6] = ew
        return table

    def apply_host(self, inputs, table):
        """numpy path for host arrays (what the reference does for every device)."""
        inputs = np.asarray(inputs)
        table = np.asarray(table)
        n, c, h, w = inputs.shape
        pad = self.pad
        out = np.empty_like(inputs)
        padded = np.pad(inputs, ((0, 0), (0, 0), (pad, pad), (pad, pad)), mode="reflect") if pad else inputs
        for i in range(n):
            cy, cx, flip, ey, ex, eh, ew = (int(v) for v in table[i, :7])
            img = padded[i, :, cy:cy + h, cx:cx + w]
            out[i] = img[:, :, ::-1] if flip else img
            if eh > 0 and ew > 0:
                out[i, :, ey:ey + eh, ex:ex + ew] = 0.0
        return np.clip(out, self.clip[0], self.clip[1]) if self.clip is not None else out

    def __call__(self, x, table):
        xb = _backend(x)
        if not isinstance(xb, BackendTensor):
            return self.apply_host(x, table)
        dev = xb.device
        tb = _backend(table)
        if not isinstance(tb, BackendTensor):
            tb = BackendTensor(np.ascontiguousarray(table, dtype=np.float32), device=dev)
        n, c, h, w = xb.shape
        if tuple(tb.shape) != (n, FIELDS):
            raise ValueError("BatchAugment: the table must have shape (%d, %d), got %s" % (n, FIELDS, tuple(tb.shape)))
        if dev.name != "cuda":  # the numpy device: the reference's own host code path
            out = BackendTensor(self.apply_host(xb.numpy(), tb.numpy()), device=dev)
        else:                   # no host fallback on the GPU: a missing extension raises in the device wrapper
            xb, tb = xb.compact(), tb.compact()
            out = BackendTensor.make((n, c, h, w), device=dev)
            lo, hi = self.clip if self.clip is not None else (0.0, 0.0)
            dev.augment_batch(xb._handle, out._handle, tb._handle, n, c, h, w, self.pad, self.clip is not None, lo, hi)
        return Tensor(out, device=dev) if isinstance(x, Tensor) else out


def smooth_one_hot(labels, num_classes, eps=0.0, device=None):
    """Dense float32 targets `onehot * (1 - eps) + eps / num_classes` (reference: test/ResNet_CIFAR10_cuda.py:181-183;
    eps = 0 gives the plain one-hot rows of the other scripts). `labels` holds class indices: a Tensor / BackendTensor on
    a device (stored as floats, as every buffer of the reference is) or a host array (then `device` says where the
    result goes; None keeps it a numpy array)."""
    on, off = np.float32(1.0 - eps), np.float32(eps / num_classes)
    lb = _backend(labels)
    if isinstance(lb, BackendTensor):
        device = lb.device
    if device is None or device.name != "cuda":  # host arrays / the numpy device; on the GPU there is no host fallback
        idx = np.asarray(lb.numpy() if isinstance(lb, BackendTensor) else labels).reshape(-1).astype(np.int64)
        hot = (idx[:, None] == np.arange(num_classes)[None, :]).astype(np.float32)
        out = hot * on + off
        if device is None:
            return out
        out = BackendTensor(out, device=device)
    else:
        if not isinstance(lb, BackendTensor):
            lb = BackendTensor(np.ascontiguousarray(labels, dtype=np.float32).reshape(-1), device=device)
        lb = lb.compact()
        n = int(np.prod(lb.shape))
        out = BackendTensor.make((n, num_classes), device=device)
        device.onehot_smooth(lb._handle, out._handle, n, num_classes, float(on), float(off))
    return out if isinstance(labels, BackendTensor) else Tensor(out, device=device)
