#!/usr/bin/env python
"""Concurrent device->host copy ceiling of this box: one plain cudaMemcpyAsync stream per GPU, 4 GB each into page-locked
host memory, all at once (one host thread per GPU). The end-to-end leg of bench.py moves 519 B of histogram per pixel to
the host; this is what the box can take at most.

    python scripts/d2h_ceiling.py [--gpus N] [--gib 4] [--reps 3]

Prints one JSON line: aggregate GB/s for 1..N GPUs copying at once, per-GPU shares, host NUMA nodes.
"""
import argparse
import json
import os
import threading
import time

import torch


def run(gpus, gib, reps):
    n = gib << 30
    src = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in gpus]
    dst = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in gpus]
    for s, d in zip(src, dst):
        d.copy_(s, non_blocking=True)
    for g in gpus:
        torch.cuda.synchronize(g)
    per = [0.0] * len(gpus)
    start = threading.Barrier(len(gpus) + 1)

    def worker(i):
        torch.cuda.set_device(gpus[i])
        start.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst[i].copy_(src[i], non_blocking=True)
        torch.cuda.synchronize(gpus[i])
        per[i] = n * reps / (time.perf_counter() - t0) / 1e9

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(len(gpus))]
    for t in ts:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    return n * reps * len(gpus) / dt / 1e9, per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--gib", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    out = {"gib_per_copy": a.gib, "reps": a.reps, "host_cores": os.cpu_count(), "levels": []}
    k = 1
    while k <= a.gpus:
        total, per = run(list(range(k)), a.gib, a.reps)
        out["levels"].append({"gpus": k, "aggregate_gb_per_s": total, "per_gpu_gb_per_s": per})
        k *= 2
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except OSError:
        out["numa_nodes"] = None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
