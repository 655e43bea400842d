"""Data-parallel replicas (new: the reference has no dist/ package).

One process per GPU (the autograd tape and the grad switch are process-global). Rank r trains on
its shard of the global batch; after `backward()` the gradients of all parameters are packed into
flat buckets and sum-all-reduced with NCCL on the communication stream while the host goes on;
`Optimizer.step()` waits for the communication stream and folds the 1/world_size average into the
fused optimizer kernel's `grad_scale`. Scripts stay unchanged apart from one `dist.init(model.parameters())`
call (which reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the launcher's environment); `backward()`
and `step()` consult this module's context.

The transport is pluggable so the bucketing logic can be tested on CPU: `PeerTransport` (the default on GPUs: NCCL
for the buckets that overlap backward, the library's own one-shot kernel over NVLink peer memory - csrc/peer.cu - for
the last, exposed one; falls back to plain `NcclTransport` when the GPUs cannot map each other);
tests/test_dist_gloo.py supplies a torch.distributed (gloo) transport for the numpy device.

Gradient accumulation over several `backward()` calls before one `step()` is not supported: after a backward the
gradients alias their (already summed) buckets, and a second backward raises instead of reducing them twice.
"""
import os
import socket
import struct
import time

from .tensor import Tensor
from .backend.backend_tensor import BackendTensor

_ctx = None


class NcclTransport:
    """NCCL over NVLink through the C ABI (dfb_comm_*). In-place sum on device buffers."""

    def __init__(self, device, rank, world, master_addr, master_port):
        self.device, self.rank, self.world = device, rank, world
        _prefer_bundled_nccl()
        uid = _exchange_unique_id(device, rank, world, master_addr, master_port)
        device.comm_init(uid, rank, world)

    def allreduce_sum(self, flat: BackendTensor):
        self.device.comm_allreduce_async((flat._handle, flat._offset), flat.size)

    def broadcast(self, flat: BackendTensor, root=0):
        self.device.comm_broadcast_async((flat._handle, flat._offset), flat.size, root)

    def wait(self):
        self.device.comm_wait()

    def close(self):
        self.device.comm_destroy()


# Which buckets the peer kernels reduce (the rest: NCCL). Measured (C4, profiles/r04*): the one-shot kernel on the exposed
# bucket beats NCCL at 2 and 8 GPUs; the two-shot pull on the overlapped buckets costs the convolutions beside it more
# than NCCL's kernels do (2 GPUs +9 us per step, 8 GPUs +170 us) - so "tail" is the default, "all" / "mid" the experiments.
_PEER_PARTS = os.environ.get("DEEPFLOWS_DP_PEER_PARTS", "tail")


_ONE_SHOT_FLOATS = 65536   # capacity of one copy in the one-shot kernel's receive area (csrc/peer.cuh: kPushCapFloats)


class PeerTransport(NcclTransport):
    """Gradient buckets in NVLink peer memory (csrc/peer.cu): every bucket is a window of one arena that all ranks of
    the box map through CUDA IPC, `allreduce_sum` is one kernel on the stream that packed the bucket (no communication
    stream, no event hops), and its completion is awaited inside the fused Adam / SGD kernel (`optimizer_waits`).
    NCCL still carries the set-up (IPC handles, the self-test's verdict) and the initial broadcast, and stays the
    transport if `peer_init` fails on any rank (e.g. GPUs that cannot map each other) - all ranks decide together."""

    optimizer_waits = True

    def __init__(self, device, rank, world, master_addr, master_port):
        super().__init__(device, rank, world, master_addr, master_port)
        self.peer = False
        self._windows = {}      # arena offset -> (slot, padded size)
        self._nccl_busy = False

    def make_buckets(self, sizes, device):
        """One flat BackendTensor per bucket inside the peer arena (None: no peer memory, allocate them normally)."""
        if self.world < 2 or not device.has("peer_init") or len(sizes) > 63:
            return None
        offsets, total = [], 0
        for n in sizes:
            offsets.append(total)
            total += (n + 3) & ~3
        try:
            arena = device.peer_init(total)
        except RuntimeError as e:
            if self.rank == 0:
                import sys
                print("DeepFlows.dist: %s; gradient buckets stay on NCCL" % (e,), file=sys.stderr)
            return None
        self.peer = True
        flats = []
        for slot, (off, n) in enumerate(zip(offsets, sizes)):
            # the last bucket (the first-registered layers) completes when backward ends: nothing overlaps its reduction
            self._windows[off] = (slot, (n + 3) & ~3, slot == len(sizes) - 1 and len(sizes) > 1)
            flats.append(BackendTensor.make((n,), (1,), device, arena, off))
        self._arena = arena
        return flats

    def allreduce_sum(self, flat):
        w = self._windows.get(flat._offset) if self.peer and flat._handle is self._arena else None
        one_shot = w is not None and w[2] and w[1] <= _ONE_SHOT_FLOATS    # the last bucket, small enough for the receive area
        if w is not None and ((_PEER_PARTS == "tail" and not one_shot) or (_PEER_PARTS == "mid" and w[2])):
            w = None   # "tail" (default): only the exposed bucket on the peer kernels; "mid": only the overlapped ones
        if w is None:
            self._nccl_busy = True
            return super().allreduce_sum(flat)
        self.device.peer_allreduce_async(flat._offset, w[1], w[0], one_shot)

    def exposed(self, flat):
        """True for the bucket the one-shot kernel reduces on the compute stream itself (DataParallel._launch_bucket)."""
        w = self._windows.get(flat._offset) if self.peer and flat._handle is self._arena else None
        return bool(w is not None and w[2] and w[1] <= _ONE_SHOT_FLOATS and _PEER_PARTS != "mid")

    def broadcast(self, flat, root=0):
        self._nccl_busy = True
        super().broadcast(flat, root)

    def wait(self):
        if self.peer:
            self.device.peer_wait()
        if self._nccl_busy:
            self._nccl_busy = False
            super().wait()

    def check(self):
        """Raises if a peer stopped answering (the kernels time out and fall through instead of hanging the GPU)."""
        if self.peer and self.device.peer_status():
            raise RuntimeError("data parallel: a peer did not answer within the time-out; gradients are not reduced")


def _prefer_bundled_nccl():
    """libdfb200 dlopens NCCL (DFB_NCCL_LIB, else libnccl.so.2 from the loader path). When the Python
    environment ships its own NCCL wheel (nvidia-nccl-cu12, the build PyTorch is tested against on this
    machine) that one is preferred: on the B200 pool the system's libnccl 2.27.3 hung inside
    ncclCommInitRank in more than half of the two-rank launches, the bundled 2.28.9 never did."""
    if os.environ.get("DFB_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["DFB_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def _exchange_unique_id(device, rank, world, addr, port):
    """Rank 0 creates the NCCL id and serves it over a plain TCP socket on MASTER_PORT + 17."""
    port = int(port) + 17
    if rank == 0:
        uid = device.comm_unique_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        for _ in range(world - 1):
            conn, _peer = srv.accept()
            conn.sendall(struct.pack("!I", len(uid)) + uid)
            conn.close()
        srv.close()
        return uid
    deadline = time.time() + 120
    while True:
        try:
            conn = socket.create_connection((addr, port), timeout=5)
            break
        except OSError:
            if time.time() > deadline:
                raise
            time.sleep(0.1)
    n = struct.unpack("!I", _recv_exact(conn, 4))[0]
    uid = _recv_exact(conn, n)
    conn.close()
    return uid


def _recv_exact(conn, n):
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed while receiving the NCCL id")
        buf += chunk
    return buf


_BUCKETS_ON_SIDE = os.environ.get("DEEPFLOWS_BUCKETS_ON_SIDE", "1") != "0"   # 0: the compute stream joins the side stream before a bucket is packed


class DataParallel:
    """Bucketed gradient all-reduce for a fixed parameter list, overlapped with backward.

    Parameters are assigned to flat buckets in reverse registration order (the last layers finish backward
    first). During `backward()` every gradient contribution to a leaf is counted; when all parameters of a
    bucket have received all their contributions (one per use in the forward pass), the bucket is packed - one
    multi-tensor copy launch - and its all-reduce is enqueued on the communication stream, while backward goes
    on producing the gradients of earlier layers on the compute stream. Whatever is still open when backward
    ends (parameters without gradient) is flushed then. Inside a captured CUDA graph the same calls become
    parallel branches of the step graph."""

    def __init__(self, params, transport, bucket_mb=4.0, tail_mb=None):
        self.params = [p for p in params]
        # The bucket that completes LAST (the first-registered layers) is the only one whose reduction nothing overlaps:
        # it is kept small, so what stands between the end of backward and the optimizer is a latency, not a transfer.
        if tail_mb is None:
            tail_mb = float(os.environ.get("DEEPFLOWS_DP_TAIL_MB", "0.25"))
        self.tail_elems = int(min(tail_mb, bucket_mb) * (1 << 20) / 4)
        self.transport = transport
        self.world = transport.world
        self.rank = transport.rank
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20) / 4))
        self._pending = False
        self._plan = None       # [(flat BackendTensor, [(param index, offset, size)])]
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._bucket_of = None  # param index -> bucket index
        self._reset_step()

    def _reset_step(self):
        self._arrived = {}      # param index -> contributions seen in this backward
        self._open = None       # bucket index -> parameters still missing
        self._launched = set()

    # buckets are filled in reverse registration order: the last layers finish backward first
    def _build_plan(self):
        dev = self.params[0].device
        plan, cur, cur_n = [], [], 0
        # the first-registered parameters (as many as fit tail_elems) form the last bucket
        tail, acc = 0, 0
        for i in range(len(self.params)):
            acc += self.params[i].data.size
            if acc > self.tail_elems:
                break
            tail = i + 1
        if tail == len(self.params):
            tail = 0
        for i in reversed(range(len(self.params))):
            n = self.params[i].data.size
            if tail and i == tail - 1 and cur:
                plan.append((cur, cur_n))
                cur, cur_n = [], 0
            if cur and cur_n + n > self.bucket_elems:
                plan.append((cur, cur_n))
                cur, cur_n = [], 0
            cur.append((i, cur_n, n))
            cur_n += n
        if cur:
            plan.append((cur, cur_n))
        flats = None
        if hasattr(self.transport, "make_buckets"):
            flats = self.transport.make_buckets([n for _, n in plan], dev)
        if flats is None:
            flats = [BackendTensor.make((n,), device=dev) for _, n in plan]
        self._plan = [(flat, slots) for flat, (slots, _) in zip(flats, plan)]
        self._bucket_of = {}
        for b, (_, slots) in enumerate(self._plan):
            for i, _, _ in slots:
                self._bucket_of[i] = b

    def broadcast_parameters(self, root=0):
        """Make every replica start from rank `root`'s weights."""
        for p in self.params:
            if not p.data.is_dense():
                p.data = p.data.compact()
            self.transport.broadcast(p.data.flat_storage(), root)  # memory order; every rank has the same layout
        self.transport.wait()

    # ---- during backward ----------------------------------------------------------------------------------
    def grad_arrived(self, param):
        """Tensor._grad_ready_hook: one more gradient contribution reached leaf `param`."""
        if self.world == 1:
            return
        i = self._index.get(id(param))
        if i is None:
            return
        if self._plan is None:
            self._build_plan()
        if self._open is None:
            if self._pending:
                # a second backward() before optimizer.step(): param.grad aliases a bucket whose all-reduce may still
                # be in flight, and re-reducing the summed values would scale them by the world size
                self.transport.wait()
                self._pending = False
                raise RuntimeError("data parallel: backward() was called again before optimizer.step(); gradient "
                                   "accumulation across backward passes is not supported (call step() / zero_grad())")
            self._open = {b: len(slots) for b, (_, slots) in enumerate(self._plan)}
        seen = self._arrived.get(i, 0) + 1
        self._arrived[i] = seen
        if seen == max(1, len(param.children)):  # one contribution per use in the forward pass
            b = self._bucket_of[i]
            self._open[b] -= 1
            if self._open[b] == 0:
                self._launch_bucket(b)

    def _launch_bucket(self, b):
        flat, slots = self._plan[b]
        dev = flat.device
        # The bucket's weight gradients may still be on the side stream (lagged joins of conv backward). Packing and reducing
        # the bucket as one more side task orders it behind them WITHOUT making the compute stream wait: the side stream forks
        # from everything the compute stream has enqueued so far (the BatchNorm gradients), the all-reduce is ordered after
        # the pack through the event it records on the current (= side) stream, the optimizer waits for the communication
        # stream, and backward() joins the side stream when it ends.
        on_side = dev.has("side_join_lag") and dev.has("side_begin") and _BUCKETS_ON_SIDE
        if on_side and getattr(self.transport, "exposed", lambda f: False)(flat):
            # the last bucket, reduced by the one-shot peer kernel: nothing is left to overlap, so the compute stream joins the
            # side stream now (backward() would a moment later) and pack, reduction and optimizer follow each other on it
            on_side = False
        if on_side:
            dev.side_begin()
        elif dev.has("side_join"):
            dev.side_join()
        try:
            self._pack_and_reduce(b, flat, slots, dev)
        finally:
            if on_side:
                dev.side_end()
        self._launched.add(b)
        self._pending = True

    def _pack_and_reduce(self, b, flat, slots, dev):
        srcs, dsts, sizes = [], [], []
        for i, off, n in slots:
            p = self.params[i]
            g = p.grad
            if not p.data.is_dense():
                p.data = p.data.compact()
            if g is None:
                flat[off:off + n] = 0.0
                continue
            g = g.with_layout_of(p.data)  # packed in the parameter's memory order (channels-last conv weights stay so)
            if dev.has("multi_copy"):
                srcs.append((g._handle, g._offset))
                dsts.append((flat._handle, flat._offset + off))
                sizes.append(n)
            else:
                flat[off:off + n] = g.flat_storage()
        if sizes:
            dev.multi_copy(srcs, dsts, sizes)
        self.transport.allreduce_sum(flat)
        # gradients now alias the bucket: the optimizer reads the reduced values in place (with_layout_of accepts a
        # dense view at an offset, the fused steps take (handle, offset) pairs)
        for i, off, n in slots:
            p = self.params[i]
            if p.grad is not None:
                p.grad = BackendTensor.make(p.data.shape, p.data.strides, p.device, flat._handle, flat._offset + off)

    # ---- after backward -----------------------------------------------------------------------------------
    def reduce_gradients(self):
        """Tensor._post_backward_hook: flush the buckets that did not complete during backward."""
        if self.world == 1:
            return
        if self._plan is None:
            self._build_plan()
        for b, (flat, slots) in enumerate(self._plan):
            if b in self._launched:
                for i, off, n in slots:  # a late contribution would have replaced the bucket view
                    g = self.params[i].grad
                    if g is not None and (g._handle is not flat._handle or g._offset != flat._offset + off):
                        raise RuntimeError("data parallel: parameter %d received a gradient after its bucket was reduced" % i)
            else:
                self._launch_bucket(b)
        self._reset_step()

    def pre_step(self, fused=False):
        """Called by Optimizer.step(): order the compute stream after the reductions and return the
        gradient scale (1/world) to fold into the fused optimizer kernel. `fused`: the caller is about to launch
        multi_adam_step / multi_sgd_step, which wait for peer-memory buckets inside their kernel."""
        if self._pending:
            if not (fused and getattr(self.transport, "optimizer_waits", False) and getattr(self.transport, "peer", False)
                    and not getattr(self.transport, "_nccl_busy", False)):
                self.transport.wait()
            self._pending = False
        return 1.0 / self.world


def init(params, transport=None, bucket_mb=None, broadcast=True):
    """Enable data parallelism for `params` (usually `model.parameters()`). `bucket_mb`: size of the gradient buckets that
    are reduced while backward runs (default 4, or DEEPFLOWS_DP_BUCKET_MB)."""
    global _ctx
    params = list(params)
    if bucket_mb is None:
        bucket_mb = float(os.environ.get("DEEPFLOWS_DP_BUCKET_MB", "4"))
    if transport is None:
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        dev = params[0].device
        # DEEPFLOWS_DP_TRANSPORT=nccl keeps the buckets on ncclAllReduce (the round-1 path; before / after numbers)
        cls = NcclTransport if os.environ.get("DEEPFLOWS_DP_TRANSPORT", "peer") == "nccl" else PeerTransport
        transport = cls(dev, rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"),
                        os.environ.get("MASTER_PORT", "29500"))
    _ctx = DataParallel(params, transport, bucket_mb)
    Tensor._post_backward_hook = _ctx.reduce_gradients
    Tensor._grad_ready_hook = _ctx.grad_arrived
    if broadcast and _ctx.world > 1:
        _ctx.broadcast_parameters(0)
    return _ctx


def shutdown():
    global _ctx
    if _ctx is not None:
        Tensor._post_backward_hook = None
        Tensor._grad_ready_hook = None
        try:
            _ctx.transport.close()
        finally:
            _ctx = None


def context():
    return _ctx


def pre_step(fused=False):
    return _ctx.pre_step(fused) if _ctx is not None else 1.0


def get_rank():
    return _ctx.rank if _ctx is not None else int(os.environ.get("RANK", "0"))


def get_world_size():
    return _ctx.world if _ctx is not None else 1
