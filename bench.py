#!/usr/bin/env python
"""bench.py - training throughput through the unchanged DeepFlows API on deepflows_b200's `cuda` device.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4] [--precision tf32|fp32|bf16]
                    [--mode graph|eager] [--impl ours|reference]

Default workload = the headline: ResNet-18 (CIFAR-10 shape, BASELINE.json configs[3]), batch 256 per GPU, BatchNorm +
Adam, data-parallel over the N GPUs of one node. `--config c2|c3|c5-vgg16|c5-resnet` selects the other BASELINE configs
(CNN-MNIST, CNN-CIFAR10, VGG-16+BN / ResNet[3,4,6,3] at 224x224, batch 128 per GPU).

One JSON line on stdout (rank 0). A "step" = batch -> forward -> softmax-CE -> backward -> (gradient all-reduce) ->
optimizer step.
  value   : images/s with the batch already resident in HBM: K steps bracketed by barrier + synchronize and CUDA events on
            the library's compute stream, max over ranks. The K-step block is repeated until >= 1 s has been timed and the
            MEDIAN block is reported (`blocks` lists how many, `block_ms` their spread): 20 steps of 1 ms on a cold GPU at
            boost clocks is a burst, not a throughput.
  e2e     : the same metric through the public API with HOST inputs: every step's batch is copied from pinned host memory
            (prefetched on a copy stream while the previous step computes, then moved into the step's input buffers) and
            every step's loss is read back to the host; all inside the timed region.
  roofline: the fused call that costs the step most device time, re-timed over rotating argument sets (> 256 MB of
            distinct buffers, i.e. out of L2 like inside the step), algorithmic bytes / FLOPs over that time,
            against MEASURED_PEAKS.json (HBM) or profiles/measured_tf32_peak.json (tensor pipe).
  parity  : the same model / weights / batch stepped once on the oracle's numpy device and on cuda (loss, logits).
  cpu_baseline / --impl reference : the reference algorithm (oracle numpy device under the same host code) on the host
            cores, on a bounded sample.
  extra   : the same step launched eagerly from Python, and in fp32 mode (N = 1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import deepflows_b200  # noqa: E402,F401
import workloads  # noqa: E402

F32 = np.float32
WIDTHS, LAYERS, HW, CLASSES = (32, 64, 128, 256), (2, 2, 2, 2), 32, 10

# name -> model builder, input shape, batch per GPU, optimizer, target smoothing, bounded CPU sample (batch, steps, warm-up)
CONFIGS = {
    "c2": dict(metric="CNN-MNIST train img/s",
               workload="CNN-MNIST (test/CNN_MNIST_cuda.py:72-96: 2 x (conv5x5-ReLU-pool) + FC), Adam lr 1e-3, one-hot targets",
               build=lambda df, d: workloads.cnn_mnist(df, d), shape=(1, 28, 28), batch=256, smooth=0.0,
               opt=lambda optim, ps: optim.Adam(ps, lr=1e-3), cpu=(256, 3, 1)),
    "c3": dict(metric="CNN-CIFAR10 train img/s",
               workload="CNN-CIFAR10 (test/CNN_CIFAR10_cuda.py:61-114: 3 x (conv-BN-ReLU-pool), Dropout(.5), FC), Adam lr 5e-3 wd 5e-4, "
                        "one-hot targets, host-RNG dropout masks uploaded every step",
               build=lambda df, d: workloads.cnn_cifar10(df, d), shape=(3, 32, 32), batch=256, smooth=0.0,
               opt=lambda optim, ps: optim.Adam(ps, lr=5e-3, weight_decay=5e-4), cpu=(256, 3, 1)),
    "c4": dict(metric="ResNet-18 CIFAR-10 train img/s",
               workload="ResNet-18 CIFAR-10 shape (test/ResNet_CIFAR10_cuda.py, widths 32-64-128-256, all blocks registered), "
                        "Adam lr 1e-3 wd 5e-4, label-smoothed dense targets, exact dgrad",
               build=lambda df, d: workloads.resnet_cifar(df, d, widths=WIDTHS, layers=LAYERS, num_classes=CLASSES, registered=True),
               shape=(3, HW, HW), batch=256, smooth=0.05,
               opt=lambda optim, ps: optim.Adam(ps, lr=1e-3, weight_decay=5e-4), cpu=(256, 3, 1)),
    "c5-vgg16": dict(metric="VGG-16 224x224 train img/s",
                     workload="VGG-16 + BatchNorm at 224x224 (test/VGG.py:7-138, 13 conv3x3 + FC 25088-4096-4096-10), Adam lr 1e-3 wd 1e-4, "
                              "one-hot targets, host-RNG dropout masks",
                     build=lambda df, d: workloads.vgg16_bn(df, d, img=224), shape=(3, 224, 224), batch=128, smooth=0.0,
                     opt=lambda optim, ps: optim.Adam(ps, lr=1e-3, weight_decay=1e-4), cpu=(2, 1, 0)),
    "c5-resnet": dict(metric="ResNet[3,4,6,3] 224x224 train img/s",
                      workload="ResNet basic blocks [3,4,6,3], widths 64-512, stride-1 stem, no max-pool (test/ResNet.py:24-150, what "
                               "pretrained_models.py:455-460 builds for 'ResNet-50'), all blocks registered, SGD lr 0.01 momentum 0.9 wd 5e-4",
                      build=lambda df, d: workloads.resnet_imagenet(df, d), shape=(3, 224, 224), batch=128, smooth=0.0,
                      opt=lambda optim, ps: optim.SGD(ps, lr=0.01, momentum=0.9, weight_decay=5e-4), cpu=(2, 1, 0)),
}


def synthetic_batch(batch, seed, shape=(3, HW, HW), smooth=0.05):
    rng = np.random.RandomState(seed)
    x = np.clip(rng.randn(batch, *shape), -1, 1).astype(F32)  # test/ResNet_CIFAR10_cuda.py:147
    t = np.eye(CLASSES, dtype=F32)[rng.randint(0, CLASSES, batch)]
    if smooth:
        t = (t * (1 - smooth) + smooth / CLASSES).astype(F32)  # :181-183
    return x, t


def conv_layers(batch):
    """(name, N, C, H, W, K, R, pad, stride, count) of every conv of the C4 model (scripts/opbench.py)."""
    out = [("stem", batch, 3, HW, HW, WIDTHS[0], 3, 1, 1, 1)]
    h, cin = HW // 2, WIDTHS[0]
    for si, (wd, nb) in enumerate(zip(WIDTHS, LAYERS)):
        stride = 1 if si == 0 else 2
        for b in range(nb):
            s = stride if b == 0 else 1
            out.append(("l%d.b%d.conv1" % (si + 1, b), batch, cin, h, h, wd, 3, 1, s, 1))
            h2 = (h + 2 - 3) // s + 1
            out.append(("l%d.b%d.conv2" % (si + 1, b), batch, wd, h2, h2, wd, 3, 1, 1, 1))
            if b == 0 and (s != 1 or cin != wd):
                out.append(("l%d.b%d.down" % (si + 1, b), batch, cin, h, h, wd, 1, 0, s, 1))
            h, cin = h2, wd
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # nvidia-smi takes a moment to start reporting
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def mark(self):
        """Index of the next sample: samples from here on were taken after this call."""
        return len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[first:last] or self.rows[max(0, first - 1):]  # at least the sample that straddles the region
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                power.append(float(r[3]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        busy = sorted(v for v in sm if mx is None or v > 0.3 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def build_training(device_name, cfg, precision, seed=0):
    import DeepFlows
    from DeepFlows import backend_api, nn
    df = workloads.namespace(DeepFlows)
    if device_name == "cuda":
        backend_api.set_precision(precision)
    backend_api.set_dgrad_mode("exact")
    np.random.seed(seed)
    model = cfg["build"](df, device_name)
    opt = cfg["opt"](df.optim, model.parameters())
    return df, model, opt, nn.CrossEntropyLoss()


def train_step(df, model, opt, crit, x, t):
    out = model(x)
    loss = crit(out, t)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss, out


# ------------------------------------------------------------------------------------------------
def run_cpu(cfg, batch, steps, warmup, keep_first=False):
    """The reference algorithm on the host cores: same host code on the oracle's numpy device. With `keep_first` the loss
    and logits of the first step (fresh seed-0 weights, batch seed 1) are returned for the parity block."""
    from oracle import numpy_device
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor, Graph
    backend_api.register_numpy_device(numpy_device)
    df, model, opt, crit = build_training("cpu", cfg, "fp32")
    dev = backend_api.Device("cpu")
    x, t = synthetic_batch(batch, 1, cfg["shape"], cfg["smooth"])
    times, first = [], None
    np.random.seed(1234)  # dropout masks
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, out = train_step(df, model, opt, crit, Tensor(x, device=dev), Tensor(t, device=dev))
        lv = float(loss.data.numpy()[0])
        if it == 0 and keep_first:
            first = (lv, out.data.numpy().copy())
        Graph.free_graph()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    backend_api.register_numpy_device(None)
    Graph.free_graph_all()
    sec = sum(times) / len(times)
    return {"value": batch / sec, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d steps (+%d warm-up) of the workload at batch %d on the oracle numpy device (reference algorithm: "
                      "im2col + sgemm, numpy/OpenBLAS threads)" % (steps, warmup, batch), "ms_per_step": sec * 1e3}, first


def gpu_first_step(cfg, batch, precision, dev):
    """One eager training step of a FRESH seed-0 model on cuda at the CPU sample's batch (for the parity block)."""
    from DeepFlows import backend_api
    from DeepFlows.tensor import Tensor, Graph
    df, model, opt, crit = build_training("cuda", cfg, precision)
    x, t = synthetic_batch(batch, 1, cfg["shape"], cfg["smooth"])
    np.random.seed(1234)
    loss, out = train_step(df, model, opt, crit, Tensor(x, device=dev), Tensor(t, device=dev))
    res = (float(loss.data.numpy()[0]), out.data.numpy().copy())
    Graph.free_graph_all()
    return res


def load_peaks():
    peaks = {}
    for name in ("MEASURED_PEAKS.json", os.path.join("profiles", "measured_tf32_peak.json")):
        try:
            peaks.update(json.load(open(os.path.join(ROOT, name))))
        except (OSError, ValueError):
            pass
    hbm = (peaks.get("hbm_gbs"), "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else \
        (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
    if peaks.get("tf32_tflops"):
        tc = (peaks["tf32_tflops"], "measured cuBLAS TF32 8192^3 burst (profiles/measured_tf32_peak.json)")
    elif peaks.get("bf16_tflops"):
        tc = (peaks["bf16_tflops"] / 2, "half the measured bf16 burst (MEASURED_PEAKS.json); TF32 itself not measured")
    else:
        tc = (795.0, "half the fallback bf16 1.59 PF (B200_PROFILING.md)")
    return hbm, tc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's)")
    ap.add_argument("--precision", default=os.environ.get("DEEPFLOWS_PRECISION", "tf32"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: the step is captured once and replayed as one CUDA graph; eager: every kernel launched from Python")
    ap.add_argument("--cpu-batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the eager-mode and fp32-mode extra measurements")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="repeat the K-step block until this much has been timed")
    ap.add_argument("--dropout", default=os.environ.get("DEEPFLOWS_DROPOUT", "host"), choices=["host", "device"],
                    help="where Dropout draws its masks: host = numpy like the reference (default), device = Philox kernel")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    os.environ["DEEPFLOWS_DROPOUT"] = args.dropout   # read by DeepFlows.backend.backend_tensor at import
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch or cfg["batch"]
    cpu_batch, cpu_steps, cpu_warm = cfg["cpu"]
    if args.cpu_batch:
        cpu_batch = args.cpu_batch
    config = {"workload": cfg["workload"] + ("" if args.dropout == "host" else " [dropout masks drawn on the device: --dropout device]"),
              "name": args.config,
              "batch_per_gpu": B, "global_batch": B * world,
              "image": "x".join(str(v) for v in cfg["shape"]), "parallelism": "dp%d" % world,
              "launch": "one CUDA graph per step (captured from the unchanged DeepFlows step)" if args.mode == "graph"
                        else "eager (every kernel launched from Python)",
              "l2": "per-step activation working set far exceeds the 126 MB L2 (C4: ~1.5 GB at batch 256); no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, cpu_steps)), min(args.warmup, cpu_warm)
        base, _ = run_cpu(cfg, cpu_batch, steps, warmup)
        line = {"impl": "reference", "metric": cfg["metric"], "value": base["value"], "unit": "img/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, batch_per_gpu=cpu_batch, global_batch=cpu_batch, launch="numpy on the host cores"),
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    from DeepFlows import backend_api, dist
    from DeepFlows.tensor import Tensor, Graph
    from DeepFlows.backend.backend_tensor import BackendTensor
    dev = backend_api.cuda()
    if not dev.enabled():
        raise SystemExit("CUDA_BACKEND extension is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    dev.set_device(local_rank)

    # ---- parity block (N = 1): fresh model, one eager step at the CPU sample's batch, against the oracle -------------
    parity = None
    base = None
    if not args.no_cpu_baseline and world == 1:
        got_loss, got_logits = gpu_first_step(cfg, cpu_batch, args.precision, dev)
        base, (want_loss, want_logits) = run_cpu(cfg, cpu_batch, cpu_steps, cpu_warm, keep_first=True)
        parity = {"batch": cpu_batch, "loss_gpu": got_loss, "loss_cpu": want_loss,
                  "rel": abs(got_loss - want_loss) / max(abs(want_loss), 1e-30),
                  "logits_rel": float(np.abs(got_logits - want_logits).max() / max(np.abs(want_logits).max(), 1e-30)),
                  "tolerance": 2e-2 if args.precision != "fp32" else 1e-4,
                  "what": "first training step of the same seed-0 model on the same batch: cuda (%s, eager) vs the oracle numpy device"
                          % args.precision}
        parity["ok"] = bool(parity["rel"] <= parity["tolerance"] and parity["logits_rel"] <= parity["tolerance"])
        if not parity["ok"]:
            raise SystemExit("bench: parity check failed: %s" % json.dumps(parity))

    df, model, opt, crit = build_training("cuda", cfg, args.precision)
    if world > 1:
        dist.init(model.parameters())
    x_host, t_host = synthetic_batch(B, 100 + rank, cfg["shape"], cfg["smooth"])
    px, pt = dev.pinned_empty(x_host.size), dev.pinned_empty(t_host.size)
    px[:], pt[:] = x_host.reshape(-1), t_host.reshape(-1)
    x_dev = Tensor(backend_api.Btensor(x_host, device=dev))
    t_dev = Tensor(backend_api.Btensor(t_host, device=dev))

    def barrier():
        dev.synchronize()
        if world > 1:
            flag = BackendTensor.make((1,), device=dev)
            flag.fill(1.0)
            dev.comm_allreduce_async(flag._handle, 1)
            dev.comm_wait()
            dev.synchronize()

    def eager_step():
        loss, _ = train_step(df, model, opt, crit, x_dev, t_dev)
        Graph.free_graph()
        return loss

    def make_step(mode):
        if mode == "graph":
            from DeepFlows.cuda_graph import CapturedStep
            return CapturedStep(eager_step, device=dev, warmup=1)  # call 1 eager, call 2 captures, then replays
        return eager_step

    resident_step = make_step(args.mode)
    captured = resident_step if args.mode == "graph" else None

    # End-to-end step: every step's batch travels from pinned host memory to the device and the loss travels
    # back. The copy of batch i+1 goes to a staging buffer on the copy stream while step i computes (double
    # buffering, the input pipeline of SURVEY 8f rank 2); at the start of step i+1 it is moved into the step's
    # input buffers device-to-device. All of it happens inside the timed region, once per step.
    stage_x, stage_t = dev.Array(x_host.size), dev.Array(t_host.size)

    def prefetch_batch():
        dev.prefetch_from_pinned(px, stage_x, x_host.size)
        dev.prefetch_from_pinned(pt, stage_t, t_host.size)

    loss_host = dev.pinned_empty(2)                           # the loss of step i lands in loss_host[i % 2]
    loss_ev = [dev.event_create(), dev.event_create()]
    e2e = {"i": 0, "pending": None, "last": None}

    def e2e_step():
        i = e2e["i"]
        dev.prefetch_wait()                                   # this step's batch has landed in the staging buffers
        dev.copy(stage_x, x_dev.data._handle, x_host.size)
        dev.copy(stage_t, t_dev.data._handle, t_host.size)
        prefetch_batch()                                      # next step's batch: waits for the two copies above only,
        loss = resident_step()                                # then runs beside this step's kernels
        # device -> host read of this step's result: an async copy into pinned memory behind the step; the host
        # picks it up one step later, so the read does not stall the launch of the next step
        dev.to_pinned_async(loss.data._handle, loss_host[i % 2:i % 2 + 1], 1)
        dev.event_record(loss_ev[i % 2])
        if e2e["pending"] is not None:
            j = e2e["pending"]
            dev.event_synchronize(loss_ev[j])
            e2e["last"] = float(loss_host[j])
        e2e["pending"] = i % 2
        e2e["i"] = i + 1
        return e2e["last"]

    def timed(fn, steps, cap=None):
        ev0, ev1 = dev.event_create(), dev.event_create()
        barrier()
        l0 = dev.launch_count()
        dev.event_record(ev0)
        for _ in range(steps):
            fn()
        dev.event_record(ev1)
        dev.event_synchronize(ev1)
        barrier()
        ms = dev.event_elapsed_ms(ev0, ev1)
        launches = dev.launch_count() - l0
        if cap is not None and cap.captured:
            launches += cap.node_counts()[0] * steps  # kernel nodes replayed by the graph launches
        dev.event_destroy(ev0)
        dev.event_destroy(ev1)
        return ms, launches

    def max_over_ranks(values):
        """Element-wise max over ranks of a list of floats (a sum all-reduce of a one-hot [rank][i] table)."""
        if world == 1:
            return list(values)
        v = np.zeros((world, len(values)), F32)
        v[rank] = values
        buf = backend_api.Btensor(v.reshape(-1), device=dev)
        dev.comm_allreduce_async(buf._handle, v.size)
        dev.comm_wait()
        return [float(m) for m in buf.numpy().reshape(world, -1).max(axis=0)]

    def timed_blocks(fn, steps, cap):
        """Blocks of `steps` steps until >= min_seconds has been timed (same count on every rank); median block."""
        ms0, launches = timed(fn, steps, cap)
        ms0 = max_over_ranks([ms0])[0]
        n_blocks = int(min(200, max(1, np.ceil(args.min_seconds * 1e3 / max(ms0, 1e-3)))))
        blocks = [ms0]
        for _ in range(n_blocks - 1):
            ms, _ = timed(fn, steps, cap)
            blocks.append(ms)
        blocks = max_over_ranks(blocks)
        return float(np.median(blocks)), blocks, launches

    W = max(3, args.warmup)
    for _ in range(W):
        resident_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # 10 ms samples; returns once nvidia-smi reports
    for _ in range(3):
        resident_step()      # every rank (the step holds collectives): the GPU is under load when timing starts
    s0 = sampler.mark()
    ms_res, blocks_res, launches = timed_blocks(resident_step, args.steps, captured)
    prefetch_batch()
    for _ in range(2):
        e2e_step()
    ms_e2e, blocks_e2e, _ = timed_blocks(e2e_step, args.steps, captured)
    s1 = sampler.mark()
    clocks = sampler.stop(s0, s1) if rank == 0 else None

    # ---- data-parallel check: replicas must hold bit-identical parameters after the timed steps ------------------------
    dp_check = None
    if world > 1:
        dev.synchronize()
        checksum = float(sum(float(p.data.numpy().astype(np.float64).sum()) for p in model.parameters()))
        v = np.zeros(world, np.float64)
        v[rank] = checksum
        hi, lo = v.astype(F32), (v - v.astype(F32).astype(np.float64)).astype(F32)  # two float32 words per rank
        buf = backend_api.Btensor(np.concatenate([hi, lo]), device=dev)
        dev.comm_allreduce_async(buf._handle, 2 * world)
        dev.comm_wait()
        got = buf.numpy()
        sums = got[:world].astype(np.float64) + got[world:].astype(np.float64)
        agree = bool(np.all(sums == sums[0]))
        tr = dist.context().transport
        if hasattr(tr, "check"):
            tr.check()   # a peer that stopped answering makes the kernels fall through: never report such a run
        dp_check = {"param_checksum": checksum, "ranks_agree": agree,
                    "transport": "peer-memory all-reduce kernel over NVLink (csrc/peer.cu), awaited inside multi_adam_kernel"
                    if getattr(tr, "peer", False) else "ncclAllReduce on the communication stream"}
        if not agree:
            # which parameters: every rank's per-parameter checksums side by side (same two-word transport)
            named = list(model.named_parameters())
            per = np.array([float(p.data.numpy().astype(np.float64).sum()) for _, p in named])
            tab = np.zeros((world, per.size), np.float64)
            tab[rank] = per
            hi2 = tab.astype(F32)
            lo2 = (tab - hi2.astype(np.float64)).astype(F32)
            b2 = backend_api.Btensor(np.concatenate([hi2.reshape(-1), lo2.reshape(-1)]), device=dev)
            dev.comm_allreduce_async(b2._handle, 2 * tab.size)
            dev.comm_wait()
            g2 = b2.numpy().astype(np.float64)
            full = (g2[:tab.size] + g2[tab.size:]).reshape(world, per.size)
            differ = [(named[i][0], full[:, i].tolist()) for i in range(per.size) if not np.all(full[:, i] == full[0, i])]
            raise SystemExit("bench: data-parallel replicas diverged: per-rank parameter checksums %s; %d of %d parameters differ, first: %s"
                             % (sums.tolist(), len(differ), per.size, differ[:6]))

    # ---- extras (N = 1): the same step launched eagerly, and in fp32 mode ---------------------------------------------
    extra = {}
    if args.config.startswith("c5"):
        args.no_extra = True   # a second captured graph / an eager pool of a 224x224 batch-128 step does not fit beside the first
    if world == 1 and not args.no_extra:
        if args.mode == "graph":
            for _ in range(3):
                eager_step()
            ms, _ = timed(eager_step, args.steps)
            extra["eager"] = {"ms_per_step": ms / args.steps, "value": B * args.steps / (ms / 1e3), "unit": "img/s",
                              "what": "same step, every kernel launched from Python (what an unchanged script gets)"}
        if args.precision != "fp32":
            backend_api.set_precision("fp32")
            try:
                step32 = make_step(args.mode)
                for _ in range(4):
                    step32()
                ms, _ = timed(step32, args.steps, step32 if args.mode == "graph" else None)
                extra["fp32"] = {"ms_per_step": ms / args.steps, "value": B * args.steps / (ms / 1e3), "unit": "img/s",
                                 "what": "same step in fp32 mode (1e-5 parity), launch mode as the headline"}
                if args.mode == "graph":
                    step32.destroy()
            finally:
                backend_api.set_precision(args.precision)

    # ---- per-kernel profile pass: CUDA events around every fused conv call, same steps ----------------
    # every rank runs it: the steps contain the data-parallel all-reduces, so the ranks must stay in lockstep
    hbm, tc = load_peaks()
    if captured is not None and args.config.startswith("c5"):
        captured.destroy()     # hand the graph's buffers back before the eager profile steps allocate theirs
        captured = None
    roofline, flops = profile_dominant_kernel(dev, eager_step, hbm, tc)

    if world > 1:
        barrier()
        dist.shutdown()
    if rank != 0:
        if world > 1:
            os._exit(0)
        return
    total_imgs = B * world * args.steps
    value = total_imgs / (ms_res / 1e3)
    line = {
        "metric": cfg["metric"], "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"tf32": "tf32 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate"}.get(args.precision, "f32"),
        "data": "synthetic", "config": config,
        "blocks": len(blocks_res), "block_ms": {"min": min(blocks_res), "median": ms_res, "max": max(blocks_res)},
        "e2e": {"value": total_imgs / (ms_e2e / 1e3), "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(4 * (x_host.size + t_host.size)), "d2h_bytes_per_step": 4, "blocks": len(blocks_e2e)},
        "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
        "gemm_tflops_per_gpu": flops / (ms_res / args.steps / 1e3) / 1e12 if flops else None,
        "clocks": clocks, "roofline": roofline,
    }
    if parity is not None:
        line["parity"] = parity
    if dp_check is not None:
        line["dp_check"] = dp_check
    if extra:
        line["extra"] = extra
    if base is not None:
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        # leave without running destructors that could wait on the other ranks (NCCL, captured graphs)
        sys.stdout.flush()
        os._exit(0)


def profile_dominant_kernel(dev, step_fn, hbm, tc):
    """Find the fused call + geometry that costs the step most device time and report it against its roofline.

    Pass 1 (eager, CUDA events around every conv2d_fprop / dgrad / wgrad / gemm / BatchNorm call of two identical steps)
    finds the distinct (op, geometry) pairs, how often each runs per step, and keeps one set of live arguments for each.
    Pass 2 re-times each pair the way it runs inside the captured step: R calls replayed from a CUDA graph (no host launch
    cost), each call on its OWN copy of every array argument with R chosen so that the copies exceed 256 MB - twice the L2 -
    so operands come from HBM as they do inside the step (back-to-back calls on one argument set would be served by the
    126 MB L2). The pair with the largest per-step total is the dominant kernel; `achieved` = its algorithmic bytes or FLOPs
     over that duration. Also returns the GEMM FLOPs of one step (sum over the recorded conv / gemm calls)."""
    pending, kept = [], set()
    for _ in range(2):  # un-instrumented eager steps: refill the allocator pool after the graph capture
        step_fn()
    dev.synchronize()

    def wrap(name, geom_of, orig):
        def fn(*a):
            e0, e1 = dev.event_create(), dev.event_create()
            dev.event_record(e0)
            orig(*a)
            dev.event_record(e1)
            geom = geom_of(a)
            # live arguments are kept for ONE call per distinct (op, geometry): holding all of them would pin every
            # activation of two steps (far beyond HBM for the 224x224 configs)
            first = (name, geom) not in kept
            kept.add((name, geom))
            pending.append((name, geom, e0, e1, a if first else None))
        return fn

    conv_geom = lambda off: (lambda a: tuple(int(v) for v in a[off:off + 8]))  # noqa: E731
    saved = {}
    specs = [("conv2d_fprop", conv_geom(4)), ("conv2d_dgrad", conv_geom(3)), ("conv2d_wgrad", conv_geom(4)),
             ("conv2d_fprop_stats", conv_geom(5)), ("conv2d_dgrad_fused", lambda a: tuple(int(v) for v in a[4:12]) + (int(a[14] is not None), int(a[15] is not None) + int(a[16] is not None),
                                                                              int(bool(a[18])), int(a[19] is not None))),
             ("gemm", lambda a: tuple(int(v) for v in a[3:8])),
             ("bn_fwd_train", lambda a: (int(a[10]), int(a[11]))), ("bn_bwd", lambda a: (int(a[8]), int(a[9]))),
             ("bn_fwd_apply", lambda a: (int(a[4]), int(a[5]), int(a[1] is not None), int(a[2] is not None), int(bool(a[6])))),
             ("bn_bwd_apply", lambda a: (int(a[8]), int(a[9]))), ("relu_bwd_bn", lambda a: (int(a[5]), int(a[6])))]
    for name, geom in specs:
        if not dev.has(name):
            continue
        orig = getattr(dev, name)
        saved[name] = orig
        dev.__dict__[name] = wrap(name, geom, orig)
    steps = 2
    try:
        for _ in range(steps):
            step_fn()
        dev.synchronize()
    finally:
        for name, orig in saved.items():
            dev.__dict__[name] = orig
    pairs = {}
    for name, geom, e0, e1, a in pending:
        eager_ms = dev.event_elapsed_ms(e0, e1)
        dev.event_destroy(e0)
        dev.event_destroy(e1)
        r = pairs.setdefault((name, geom), {"count": 0, "eager_ms": 0.0, "args": a})
        if r["args"] is None and a is not None:
            r["args"] = a
        r["count"] += 1
        r["eager_ms"] += eager_ms
    if not pairs:
        return None, 0.0

    def conv_work(geom):
        n, c, h, w, k, rr, p, s = geom
        oh, ow = (h + 2 * p - rr) // s + 1, (w + 2 * p - rr) // s + 1
        return 2.0 * n * oh * ow * k * c * rr * rr, 4.0 * (n * c * h * w + k * c * rr * rr + n * oh * ow * k)

    flops_step = 0.0
    for (name, geom), r in pairs.items():
        if name.startswith("conv2d"):
            flops_step += conv_work(geom[:8])[0] * r["count"] / steps
        elif name == "gemm":
            flops_step += 2.0 * geom[0] * geom[1] * geom[2] * r["count"] / steps

    is_array = lambda v: hasattr(v, "size") and hasattr(v, "ptr")  # noqa: E731  (the shim's Array)

    def clone_args(a):
        out = []
        for v in a:
            if is_array(v):
                c = dev.Array(v.size)
                dev.copy(v, c, v.size)
                out.append(c)
            elif isinstance(v, tuple) and not (len(v) == 2 and is_array(v[0]) and isinstance(v[1], int)):
                out.append(tuple(clone_args(v)))   # a BatchNorm record: (x, statistics, gamma, ...)
            else:
                out.append(v)
        return out

    def arg_floats(a):
        return sum(v.size if is_array(v) else (arg_floats(v) if isinstance(v, tuple) else 0) for v in a)

    # pass 2: graph-replayed timing of every distinct pair over rotating (L2-cold) argument sets
    for (name, geom), r in pairs.items():
        fn, a = saved[name], r["args"]
        per_set = 4 * arg_floats(a)
        reps = int(max(4, min(48, np.ceil((256 << 20) / max(per_set, 1)))))
        if per_set * reps > (8 << 30):   # the 224x224 layers: a few sets are already far beyond L2
            reps = max(2, int((8 << 30) // per_set))
        sets = [list(a)] + [clone_args(a) for _ in range(reps - 1)]
        for s_ in sets[:2]:
            fn(*s_)
        dev.graph_begin_capture()
        for s_ in sets:
            fn(*s_)
        g = dev.graph_end_capture()
        dev.graph_launch(g)
        e0, e1 = dev.event_create(), dev.event_create()
        dev.event_record(e0)
        for _ in range(3):
            dev.graph_launch(g)
        dev.event_record(e1)
        dev.event_synchronize(e1)
        r["us"] = dev.event_elapsed_ms(e0, e1) / (3 * reps) * 1e3
        r["per_step"] = r["count"] / steps
        r["sets"] = reps
        dev.event_destroy(e0)
        dev.event_destroy(e1)
        dev.graph_destroy(g)
        r["args"] = None
        del sets
    by_op = {}
    for (name, geom), r in pairs.items():
        by_op[name] = by_op.get(name, 0.0) + r["us"] * r["per_step"] / 1e3
    (name, geom), r = max(pairs.items(), key=lambda kv: kv[1]["us"] * kv[1]["per_step"])
    avg_s = r["us"] * 1e-6
    if name.startswith("conv2d") or name == "gemm":
        if name == "gemm":
            m_, n_, k_ = geom[:3]
            flops, bytes_min = 2.0 * m_ * n_ * k_, 4.0 * (m_ * k_ + k_ * n_ + m_ * n_)
            desc = "gemm M=%d N=%d K=%d ta=%d tb=%d" % geom
        else:
            flops, bytes_min = conv_work(geom[:8])
            if name == "conv2d_dgrad_fused":   # + the addend, the BatchNorm inputs and the residual it reads, all of the output's size
                n_, c_, h_, w_ = geom[:4]
                bytes_min += 4.0 * n_ * c_ * h_ * w_ * (geom[8] + geom[9] + geom[11])
            desc = "%s N=%d C=%d H=%d W=%d K=%d R=%d pad=%d stride=%d" % ((name,) + geom[:8])
            if len(geom) > 8:
                desc += " addend=%d bn=%d relu=%d res=%d" % geom[8:12]
        t_tc_us = flops / (tc[0] * 1e12) * 1e6
        t_hbm_us = bytes_min / (hbm[0] * 1e9) * 1e6
        bound = "tensor" if t_tc_us > t_hbm_us else "hbm"
        achieved = flops / avg_s / 1e12 if bound == "tensor" else bytes_min / avg_s / 1e9
        unit = "TFLOP/s" if bound == "tensor" else "GB/s"
        extra = {"flops_per_launch": flops, "bytes_per_launch": bytes_min, "tflops": flops / avg_s / 1e12,
                 "gbs": bytes_min / avg_s / 1e9, "t_hbm_us": t_hbm_us, "t_tensor_us": t_tc_us}
    else:
        rows, c = geom[:2]
        # bytes per element: two-pass BatchNorm forward 12, backward 20; the one-pass halves of the fused
        # path: apply = read each input + write (8 + 4 per extra input), backward apply / ReLU-through-BN 12 (+ 4 per extra)
        per_elem = {"bn_fwd_train": 12.0, "bn_bwd": 20.0, "bn_bwd_apply": 12.0}.get(name)
        if name == "bn_fwd_apply":
            per_elem = 8.0 + 4.0 * (geom[2] + geom[3])
        if name == "relu_bwd_bn":
            per_elem = 12.0
        bytes_min = per_elem * rows * c
        bound, achieved, unit = "hbm", bytes_min / avg_s / 1e9, "GB/s"
        desc = "%s rows=%d C=%d" % (name, rows, c) + ("" if len(geom) == 2 else " " + ",".join(str(v) for v in geom[2:]))
        extra = {"bytes_per_launch": bytes_min}
    peak, which = (tc if bound == "tensor" else hbm)
    traffic = None
    try:  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture, when there is one
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = table.get(desc, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    worst = sorted(((kv[1]["us"] * kv[1]["per_step"], kv[0][0], kv[0][1], kv[1]["us"]) for kv in pairs.items()), reverse=True)[:12]
    return dict({"kernel": desc, "bound": bound, "achieved": achieved, "peak": peak, "peak_source": which, "unit": unit,
                 "frac": achieved / peak, "traffic": traffic, "avg_launch_us": r["us"], "launches_per_step": r["per_step"],
                 "argument_sets": r["sets"],
                 "share_of_step_ms": {k: round(v, 4) for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])},
                 "top_calls": [{"us_per_step": round(t, 2), "op": n_, "geom": list(g_), "us": round(u, 2)} for t, n_, g_, u in worst],
                 "how": "each distinct fused call of the step (found with CUDA events around the calls of eager steps) is "
                        "re-timed as R calls replayed from one CUDA graph, each call on its own copy of the arguments (R sets "
                        "> 256 MB, so nothing is served from L2), CUDA events on the compute stream around 3 replays"},
                **extra), flops_step


if __name__ == "__main__":
    main()
