#!/usr/bin/env python
"""Timeline of ONE training step replayed from its CUDA graph (bench.py's workload and step): when each kernel's work
starts (%globaltimer, recorded by the kernels themselves: dfb_trace_begin / dfb_trace_end), matched with the kernel names
the host noted while the step was captured. Start-to-start intervals along the main stream are what each kernel costs
the step (its own time plus the dependency gap behind it); side-stream kernels (the weight gradients) are listed in place.

    python scripts/step_timeline.py [--config c4] [--precision tf32] [--env K=V ...] > profiles/r02_timeline_c4.txt
"""
import argparse
import collections
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepflows_b200  # noqa: E402,F401
import bench  # noqa: E402


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("dfb::", "").replace("(anonymous namespace)::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--replays", type=int, default=3)
    args = ap.parse_args()
    from DeepFlows import backend_api
    from DeepFlows.tensor import Graph, Tensor
    from DeepFlows.cuda_graph import CapturedStep
    cfg = dict(bench.CONFIGS[args.config])
    B = args.batch or cfg["batch"]
    cfg["batch"] = B
    dev = backend_api.cuda()
    m = dev.mod
    # under torchrun (WORLD_SIZE > 1): the data-parallel step; every rank runs it, rank 0 prints its own timeline
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    backend_api.set_precision(args.precision)
    backend_api.set_dgrad_mode("exact")
    df, model, opt, crit = bench.build_training("cuda", cfg, args.precision)
    if world > 1:
        from DeepFlows import dist
        dist.init(model.parameters())
    x_host, t_host = bench.synthetic_batch(B, 100 + rank, cfg["shape"], cfg["smooth"])
    x_dev = Tensor(backend_api.Btensor(x_host, device=dev))
    t_dev = Tensor(backend_api.Btensor(t_host, device=dev))

    def eager_step():
        loss, _ = bench.train_step(df, model, opt, crit, x_dev, t_dev)
        Graph.free_graph()
        return loss

    eager_step()
    dev.synchronize()
    m.trace_begin(1 << 14)
    step = CapturedStep(eager_step, device=dev, warmup=0)
    step()                      # captures (the host notes every launch of the capture pass)
    if not step.captured:
        step()
    dev.synchronize()
    _, _, host_lines = m.trace_end(1 << 14)
    for _ in range(5):
        step()
    dev.synchronize()
    m.trace_begin(1 << 14)
    host0 = None
    for _ in range(args.replays):
        step()
    dev.synchronize()
    rec, n, _ = m.trace_end(1 << 14)
    if world > 1:
        from DeepFlows import dist
        dist.shutdown()
        prefix = os.environ.get("DFB_TIMELINE_PREFIX")   # every rank writes <prefix>_rank<r>.txt (else only rank 0 prints)
        if prefix:
            sys.stdout = open("%s_rank%d.txt" % (prefix, rank), "w")
        elif rank != 0:
            os._exit(0)
    rec = np.asarray(rec)[:n]
    per = n // args.replays
    rec = rec[(args.replays - 1) * per:]            # the last replay
    # host launches of the capture pass: "stream gx gy gz block name"
    host = []
    for line in host_lines:
        st, gx, gy, gz, bx, name = line.split(" ", 5)
        host.append((st if st != "other" else "comm", (int(gx), int(gy), int(gz), int(bx)), short(name)))
    host = host[-per:] if len(host) >= per else host
    by_fp = collections.defaultdict(collections.deque)
    for h in host:
        by_fp[h[1]].append(h)
    order = np.argsort(rec[:, 0], kind="stable")
    t0 = int(rec[order[0], 0])
    rows = []
    for i in order:
        t, fp = int(rec[i, 0]), int(rec[i, 1])
        key = (fp & 0xFFFFFF, (fp >> 24) & 0xFFFF, (fp >> 40) & 0xFFF, fp >> 52)
        h = by_fp[key].popleft() if by_fp[key] else ("?", key, "?")
        rows.append(((t - t0) / 1000.0, h[0], key, h[2]))
    print("# %s, batch %d, %s: one replay of the captured step, %d kernels (host noted %d launches in the capture pass)"
          % (args.config, B, args.precision, len(rows), len(host)))
    print("# rank %d of %d; %%globaltimer of the step's first kernel: %d ns" % (rank, world, t0))
    print("# start = when the kernel's work begins (after griddepcontrol.wait), us since the step's first kernel;")
    print("# cost = start of the next MAIN-stream kernel minus this start (main-stream kernels only)")
    print("%9s %8s %5s %-22s %s" % ("start us", "cost us", "strm", "grid x block", "kernel"))
    main_idx = [i for i, r in enumerate(rows) if r[1] not in ("side", "comm")]
    nxt = {a: b for a, b in zip(main_idx, main_idx[1:])}
    cost_by_kernel = collections.OrderedDict()
    for i, (t, st, key, name) in enumerate(rows):
        cost = rows[nxt[i]][0] - t if i in nxt else float("nan")
        print("%9.2f %8.2f %5s %-22s %s" % (t, cost, st, "%dx%dx%d x %d" % key, name[:110]))
        if st not in ("side", "comm") and i in nxt:
            a = cost_by_kernel.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += cost
    total = rows[-1][0]
    print("\n# main-stream cost by kernel (sum of start-to-start intervals; total %.1f us to the start of the last kernel)" % total)
    for k, (c, v) in sorted(cost_by_kernel.items(), key=lambda kv: -kv[1][1]):
        print("%9.1f us %5.1f%%  n=%3d  avg %6.2f  %s" % (v, 100 * v / total, c, v / c, k[:120]))


if __name__ == "__main__":
    main()
    sys.stdout.flush()
    os._exit(0)
