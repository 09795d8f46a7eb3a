#!/usr/bin/env python
"""bench.py -- filtered Mpixels/s of the single histogram filter on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3] [--dist uniform|blocky]

A "step" is one pass of the filter over one batch of synthetic chunk neighbourhoods. The default workload is C3, the
configuration BASELINE.json quotes its scaling target on: 256 neighbourhoods (3x3 chunks of 512x512 uint16 biome ids),
radius 64, 64 biomes; biome ids iid uniform like the reference's own benchmark (STPTestHistogram.cpp:342-343). With N > 1
(torchrun, one rank per GPU) the ONE 256-chunk batch is sharded contiguously, 256/N chunks per rank, the way the
reference's filterDistributed splits one job over its workers (SHF.cpp:789-868): independent units, halos replicated,
no collective on the data path, STRONG scaling (`--scaling weak` gives every rank its own full batch instead);
`value` = pixels of all ranks / max-over-ranks device time.

Printed keys beyond the base contract:
  roofline      the emitting kernel (emit_kernel<K>): algorithmic bytes of the step / its CUDA-event duration vs the measured HBM peak
  phases_ms     median CUDA-event time of every kernel phase of a step
  cpu_baseline  the reference's own filter (oracle/_ref, else the C restatement) timed on this box's host cores on a
                bounded sample of the same chunks (N=1 only)
  e2e           same metric through shf_run_batch with HOST buffers: H2D of the inputs and D2H of bins + offsets inside
                the timed region, sub-batches on two host threads so copies overlap compute
  parity_checked  sampled chunks of the TIMED result (downloaded after the timed region) compared bit for bit with the
                reference CPU filter (oracle/_ref, else the C restatement)
  consumer      the device-side heightfield consumer on the resident result (filter -> heightfield, no host round trip)
  config4_chain BASELINE.json config 4 on its own shape: one 2048x2048 chunk, filter -> heightfield on device (N=1 only)
`--impl reference` times the reference CPU filter alone (all host cores) on the same workload definition.
"""
from __future__ import annotations

import argparse
import ctypes
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "filtered_mpixels_per_s"
UNIT = "Mpixels/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "blocky", "rare", "stripes"])
    ap.add_argument("--chunks", type=int, default=0, help="override the number of chunks per GPU (debugging)")
    ap.add_argument("--radius", type=int, default=0, help="override the radius (BASELINE.json config 5 sweeps)")
    ap.add_argument("--biomes", type=int, default=0, help="override the biome count (BASELINE.json config 5 sweeps)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (BASELINE.json config 3): one batch is sharded over the GPUs; weak: every GPU filters "
                         "its own batch of the workload's size")
    ap.add_argument("--min-seconds", type=float, default=0.0,
                    help="raise --steps so that the timed region lasts at least this long (single-call workloads: the "
                         "clock sampler needs a few 50 ms samples inside it)")
    ap.add_argument("--parity-chunks", type=int, default=4, help="chunks of the timed result checked against the CPU filter")
    ap.add_argument("--no-consumer", action="store_true")
    ap.add_argument("--sync-steps", action="store_true",
                    help="time the host-blocking shf_run_device instead of back-to-back shf_run_device_async calls")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-sub", type=int, default=32, help="chunks per shf_run_batch call in the end-to-end leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--heightfield", action="store_true", help="(kept for old scripts: the consumer leg now always runs)")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"],
                    help="torch.distributed backend of the barrier / max-over-ranks plumbing (no data-path collective)")
    return ap.parse_args()


def workload_of(args):
    from superterrainplus_b200 import workloads

    wl = dataclasses.replace(workloads.CONFIGS[args.workload], dist=args.dist)
    if args.chunks:
        wl = dataclasses.replace(wl, chunks=args.chunks)
    if args.radius:
        wl = dataclasses.replace(wl, radius=args.radius)
    if args.biomes:
        wl = dataclasses.replace(wl, biomes=args.biomes)
    return wl


def shard_of(rank, world, chunks, scaling):
    """(first chunk id, chunk count) this rank filters. weak: every rank its own `chunks`-chunk batch (different chunk
    ids); strong: one `chunks`-chunk batch split into contiguous shards."""
    if scaling == "weak":
        return rank * chunks, chunks
    per = (chunks + world - 1) // world
    first = min(rank * per, chunks)
    return first, min(per, chunks - first)


def outputs_exceed_l2(wl, per_gpu):
    """Whether the histograms one GPU writes per step are certainly larger than the 126 MB L2, from the workload's
    definition alone (both arms print it, the GPU arm times accordingly): one offset per pixel plus at least one bin --
    on iid uniform ids at least half of min(biomes, window cells) bins."""
    span = 2 * wl.radius + 1
    bins = max(1, min(wl.biomes, span * span) // 2) if wl.dist == "uniform" else 1
    return per_gpu * wl.map_size[0] * wl.map_size[1] * (4 + 8 * bins) > (126 << 20)


def config_of(wl, n_gpus, scaling):
    """The same dictionary for both arms (`wl` = the workload as named, before sharding)."""
    per_gpu = wl.chunks if scaling == "weak" else (wl.chunks + n_gpus - 1) // n_gpus
    return {
        "workload": f"{wl.name}: {wl.chunks} x (3x3 neighbourhood of {wl.map_size[0]}x{wl.map_size[1]} uint16 maps), "
                    f"radius {wl.radius}, {wl.biomes} biomes, {wl.dist} ids",
        "chunks_total": wl.chunks * (n_gpus if scaling == "weak" else 1), "chunks_per_gpu": per_gpu,
        "map": list(wl.map_size), "radius": wl.radius, "biomes": wl.biomes, "distribution": wl.dist,
        "sharding": (f"one {wl.chunks}-chunk batch split into {n_gpus} contiguous shards" if scaling == "strong"
                     else f"{n_gpus} x {wl.chunks}-chunk batches") + ", halos replicated, no collective",
        "l2": ("outputs per step exceed the 126 MB L2 (no flush needed)" if outputs_exceed_l2(wl, per_gpu) else
               "outputs per step may fit the 126 MB L2: a 256 MB buffer is rewritten between steps, every step timed on its own"),
    }


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled every 20 ms on a thread (NVML through nvidia_ml_py; nvidia-smi
    polled every 50 ms when NVML cannot be loaded). Started ahead of the warm-up; stop(t0, t1) summarises the samples
    stamped inside the timed region."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []          # (time.time(), sm MHz, max sm MHz, [reasons])
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_of = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                    (getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                    (getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                    (getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(reasons_of(h))
                        self.rows.append((time.time(), mhz, max_sm, [n for b, n in bits if mask & b]))
                    except Exception:  # noqa: BLE001
                        pass
                    self._stop.wait(0.02)

            self.source = "nvml, 20 ms"
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            return
        except Exception:  # noqa: BLE001
            pass
        self.source = "nvidia-smi -lms 50"
        q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            cmd = ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"]
            if subprocess.run(["which", "stdbuf"], capture_output=True).returncode == 0:
                cmd = ["stdbuf", "-oL"] + cmd   # (a pipe would otherwise deliver the lines 4 KB at a time)
            proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.source = None
            return

        def pump():
            import datetime

            for ln in proc.stdout:
                f = [x.strip() for x in ln.split(",")]
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    self.rows.append((ts, float(f[1]), float(f[2]),
                                      [n for n, v in zip(self.NAMES, f[3:7]) if v.lower().startswith("active")]))
                except (ValueError, IndexError):
                    continue
                if self._stop.is_set():
                    proc.terminate()
                    break

        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()

    def stop(self, t0=None, t1=None):
        """Summary of the samples stamped inside [t0, t1] (time.time() values); all samples when none falls inside."""
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (nvml and nvidia-smi unavailable)"]}
        rows = list(self.rows)
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        used = inside or rows
        reasons = sorted({n for r in used for n in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in used) if used else None,
                "sm_max_mhz": max(r[2] for r in used) if used else None, "reasons": reasons,
                "samples": len(used), "samples_in_timed_region": len(inside), "source": self.source}


def measured_traffic(kernel, wl, n_chunks):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json, per chunk), scaled to the chunks of this launch; None when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f)
        for row in table:
            if (row["kernel"] == kernel and row["workload"] == wl.name and row["distribution"] == wl.dist
                    and row["radius"] == wl.radius and row["biomes"] == wl.biomes):
                return row["dram_bytes_per_chunk"] * n_chunks
    except Exception:
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own filter (oracle/_ref) or the C restatement. Used by cpu_baseline and --impl reference only.
# ----------------------------------------------------------------------------------------------------------------------
def cpu_filter_throughput(wl, host_maps, seconds, all_cores):
    """Mpixels/s of the reference CPU filter over `host_maps` (list of uint16 arrays), bounded by `seconds`."""
    import oracle

    w, h = wl.map_size
    cores = os.cpu_count() or 1
    if oracle.have_reference():
        kind = "reference"
        n_workers = max(1, cores // 4) if all_cores else 1  # every filter object owns a fixed 4-thread pool (SHF.cpp:36)
        sessions = [oracle.ReferenceSession(0xFF) for _ in range(n_workers)]
        threads_used = 4 * n_workers

        def run(sess, m):
            sess.run_raw(m, wl.map_size, wl.nn, wl.total, wl.radius)
    else:
        kind = "port"
        n_workers = cores if all_cores else 1
        sessions = [None] * n_workers
        threads_used = n_workers

        def run(sess, m):
            oracle.run_port(m, wl.map_size, wl.nn, wl.radius)

    for s in sessions:  # warm-up: the reference buffer is an adaptive pool (SHF.h:33-34)
        run(s, host_maps[0])
    done = [0] * n_workers
    t0 = time.perf_counter()
    deadline = t0 + seconds

    def worker(i):
        k = i
        while time.perf_counter() < deadline:
            run(sessions[i], host_maps[k % len(host_maps)])
            done[i] += 1
            k += n_workers

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(n_workers)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    chunks = sum(done)
    for s in sessions:
        if s is not None:
            s.close()
    return {"seconds": dt, "value": chunks * w * h / dt / 1e6, "unit": UNIT, "cores": threads_used, "kind": kind,
            "sample": f"{chunks} chunk calls over {len(host_maps)} distinct chunks of the workload in {dt:.1f} s, "
                      + (f"{n_workers} filter object(s) x 4 pool threads" if kind == "reference" else
                         f"{n_workers} thread(s) of the single-threaded C restatement")
                      + f", host has {cores} logical cores"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from superterrainplus_b200 import workloads

    wl = workload_of(args)
    n_distinct = 4
    host_maps = [workloads.make_map_np(wl, i) for i in range(n_distinct)]
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_filter_throughput(wl, host_maps, per_step, all_cores=True)
        if i >= args.warmup:
            vals.append(last["value"])
            secs.append(last["seconds"])
    value = statistics.mean(vals)
    per_step = statistics.mean(secs)
    w, h = wl.map_size
    last["value"] = value
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u16 samples, u32 counts, f32 weights", "data": "synthetic", "impl": "reference",
        "config": config_of(wl, args.gpus, args.scaling), "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """One rank per GPU: run this rank's host threads (and first-touch its page-locked buffers) on the NUMA node its GPU
    hangs off, so that the end-to-end leg's 34 GB of histograms per rank and step do not cross the socket interconnect.
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local).pci_bus_id  # available in recent torch; else nvidia-smi below
    except Exception:
        bus = None
    try:
        if not isinstance(bus, str) or ":" not in bus:
            q = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                               capture_output=True, text=True, timeout=20).stdout.strip()
            bus = q
        dom, b, rest = bus.lower().split(":")
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{b}:{rest}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None



def download_chunk(buf, chunk, w, h):
    """(items, weights, offsets) of one chunk of the buffer's device-resident result, copied to the host."""
    import torch

    from superterrainplus_b200 import api

    bins_p, offs_p = buf.readDevice()
    base = buf.chunkBase()
    lo, hi = int(base[chunk]), int(base[chunk + 1])
    per = w * h + 1
    raw = torch.as_tensor(api.DeviceArrayView(bins_p + lo * 8, 2 * (hi - lo)), device="cuda").cpu().numpy()
    offs = torch.as_tensor(api.DeviceArrayView(offs_p + chunk * per * 4, per), device="cuda").cpu().numpy()
    bins = raw.view(api.BIN_DTYPE)
    return bins["Item"].copy(), bins["Weight"].copy(), offs.view(np.uint32).copy()


def check_against_cpu_filter(wl, host_map, got):
    """Bit-exact comparison of one chunk with the reference CPU filter (the oracle as the checker, never as the product)."""
    import oracle

    if oracle.have_reference():
        sess = oracle.ReferenceSession(0xFF)
        items, weights, offsets = sess.run(host_map, wl.map_size, wl.nn, wl.radius, wl.total)
        sess.close()
        kind = "reference"
    else:
        items, weights, offsets = oracle.run_port(host_map, wl.map_size, wl.nn, wl.radius)
        kind = "port"
    ok = (np.array_equal(got[2], offsets) and np.array_equal(got[0], items)
          and np.array_equal(got[1].view(np.uint32), weights.view(np.uint32)))
    return bool(ok), kind


def heightfield_generator(pkg, api, filt, biomes):
    rng = np.random.default_rng(7)
    table = np.zeros(biomes, dtype=api.BIOME_PROPERTY_DTYPE)
    table["Scale"], table["Octave"] = rng.uniform(100.0, 900.0, biomes), 8
    table["Persistence"], table["Lacunarity"] = 0.5, 2.0
    table["Depth"], table["Variation"] = rng.uniform(0.0, 1.0, biomes), rng.uniform(0.1, 1.0, biomes)
    perm = np.tile(rng.permutation(256).astype(np.uint8), 2)
    ang = np.arange(12) * (2 * np.pi / 12)
    return pkg.STPMultiBiomeHeightfield(filt, table, perm, np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32))


def config4_chain(pkg, api, workloads, filt, local, stream, steps=8):
    """BASELINE.json config 4 on its own shape: one 2048x2048 chunk (3x3 neighbourhood, radius 64, 32 biomes) filtered and
    fed to the multi-biome heightfield kernel where the filter left the histogram in HBM -- no host round trip of bins or
    offsets (the copy + sync of STPBiomefieldGenerator.cpp:108-123 is gone). Device time per chunk, CUDA events."""
    import torch

    out = {}
    FB = pkg.STPSingleHistogramFilter.STPFilterBuffer
    for dist in ("uniform", "blocky"):
        wl = dataclasses.replace(workloads.CONFIGS["C4"], dist=dist)
        w, h = wl.map_size
        tw, th = wl.total
        info = pkg.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total)
        maps = workloads.make_maps_torch(wl, 0, 1, torch.device("cuda", local))
        buf = FB(FB.STPExecutionType.Parallel)
        gen = heightfield_generator(pkg, api, filt, wl.biomes)
        heights = torch.empty((1, h, w), dtype=torch.float32, device="cuda")
        offs = np.zeros((1, 2), dtype=np.float32)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for i in range(3 + steps):
            flush.fill_(i & 0xFF)
            k = i - 3
            if k >= 0:
                ev[k][0].record()
            filt.runDeviceAsync(maps.data_ptr(), th * tw, 1, info, buf, wl.radius, stream)
            if k >= 0:
                ev[k][1].record()
            gen(buf, 0, 1, offs, heights.data_ptr(), stream)   # completes the filter call, then enqueues the consumer
            if k >= 0:
                ev[k][2].record()
        torch.cuda.synchronize()
        f_ms = statistics.median(e[0].elapsed_time(e[1]) for e in ev)
        c_ms = statistics.median(e[0].elapsed_time(e[2]) for e in ev)
        n_bins, _ = buf.size()
        out[dist] = {"filter_ms": f_ms, "chain_ms": c_ms, "chain_mpixels_per_s": w * h / (c_ms * 1e-3) / 1e6,
                     "bins_per_pixel": n_bins / (w * h), "heights_finite": bool(torch.isfinite(heights).all().item())}
        gen.close()
        buf.close()
        del maps, heights, flush
    out["what"] = ("one 2048x2048 chunk, radius 64, 32 biomes: shf_run_device_async -> shf_heightfield_run on the "
                   "device-resident histogram, 8 octaves per bin; L2 flushed between iterations")
    return out


def d2h_ceiling(barrier, sum_over_ranks, gib=2, reps=3):
    """What the box can take: every rank copies `gib` GiB from its GPU into page-locked host memory with plain
    cudaMemcpyAsync, all ranks at once (barrier before and after), `reps` times. Returns (this job's aggregate GB/s, the
    slowest... per-rank GB/s). The end-to-end leg moves 34 GB of histograms per 256 chunks, so this is its ceiling."""
    import torch

    n = gib << 30
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt_local = time.perf_counter() - t0
    barrier()
    dt_all = time.perf_counter() - t0
    total = sum_over_ranks(float(n * reps))
    del src, dst
    return total / dt_all / 1e9, n * reps / dt_local / 1e9


def run_ours(args):
    import torch
    import torch.distributed as dist

    import superterrainplus_b200 as pkg
    from superterrainplus_b200 import api, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the filter has no CPU path (use --impl reference for the CPU filter)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda" if args.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda" if args.backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl_named = workload_of(args)
    wl = wl_named
    w, h = wl.map_size
    tw, th = wl.total
    first_chunk, n = shard_of(rank, world, wl.chunks, args.scaling)
    total_chunks = wl.chunks * world if args.scaling == "weak" else wl.chunks
    if n == 0:
        raise SystemExit("bench.py: more ranks than chunks")
    wl = dataclasses.replace(wl, chunks=n)
    info = pkg.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total)
    maps = workloads.make_maps_torch(wl, first_chunk, n, torch.device("cuda", local))
    torch.cuda.synchronize()

    filt = pkg.STPSingleHistogramFilter(local)
    FB = pkg.STPSingleHistogramFilter.STPFilterBuffer
    buf = FB(FB.STPExecutionType.Parallel)
    stream = torch.cuda.current_stream().cuda_stream
    # (the timed loop runs without the library's phase events; a second, untimed loop with them gives phases_ms and the
    # emitting kernel's duration for the roofline: seven event records per call are not free on a 1 ms step)
    api.set_profiling(False)

    # A step = one pass of the filter over this rank's shard, inputs resident in HBM. Steps are enqueued back to back
    # (shf_run_device_async); the plan checks of the last one are made by buf.wait() inside the timed region.
    def step():
        if args.sync_steps:
            filt.runDevice(maps.data_ptr(), th * tw, n, info, buf, wl.radius, stream)
        else:
            filt.runDeviceAsync(maps.data_ptr(), th * tw, n, info, buf, wl.radius, stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        step()
    repeated0 = buf.wait()
    torch.cuda.synchronize()
    est = (time.perf_counter() - t_w0) / max(1, args.warmup)
    n_bins, _ = buf.size()
    plan = buf.lastPlan()
    alg_bytes = workloads.algorithmic_bytes(wl, n_bins)
    steps = args.steps
    if args.min_seconds > 0.0:
        # (the warm-up steps grow the buffers: their wall time overstates a step, so a few more are timed for the estimate)
        t_c0 = time.perf_counter()
        for _ in range(5):
            step()
        buf.wait()
        torch.cuda.synchronize()
        est = min(est, (time.perf_counter() - t_c0) / 5)
        steps = max(steps, int(args.min_seconds / max(est, 1e-5)) + 1)
        steps = int(max_over_ranks(float(steps)))
    # outputs that fit the L2: rewrite a buffer larger than the L2 between steps and time every step on its own
    # (the rule config_of states: decided from the workload's definition, so that the line says what was done)
    flush = None if outputs_exceed_l2(wl_named, n) else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    api.stats_reset()
    barrier()
    t_wall0 = time.time()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        buf.wait()
        e1.record()
        barrier()
        ms_local = e0.elapsed_time(e1) / steps
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.fill_(i & 0xFF)
            evs[i][0].record()
            step()
            evs[i][1].record()
        buf.wait()
        barrier()
        ms_local = sum(a.elapsed_time(b_) for a, b_ in evs) / steps
    t_wall1 = time.time()
    launches, _, _ = api.stats()   # kernels launched by the K timed steps (the profiled loop below is not counted)
    repeated = buf.wait() - repeated0
    ms_step = max_over_ranks(ms_local)
    # ---- the same steps once more with the library's phase events (CUDA events on the launching stream) ----
    api.set_profiling(True)
    n_prof = min(steps, 32)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step()
    buf.wait()
    p0.record()
    for i in range(n_prof):
        if flush is not None:
            flush.fill_(i & 0xFF)
        step()
    buf.wait()
    p1.record()
    torch.cuda.synchronize()
    phases = [buf.phaseMs(back) for back in range(n_prof)]
    ms_profiled = p0.elapsed_time(p1) / n_prof if flush is None else None
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    api.set_profiling(False)
    value = total_chunks * w * h / (ms_step * 1e-3) / 1e6

    phases_ms = {k: statistics.median(p[k] for p in phases) for k in phases[0]}
    peak, peak_src = measured_peak()
    emit_ms = phases_ms["emit"]
    achieved = alg_bytes / (emit_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "emit_kernel<K>" if plan["k_sets"] else "march_generic_kernel<emit>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic("emit_kernel", wl, n), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": emit_ms,
                "whole_step_frac": alg_bytes / (ms_local * 1e-3) / 1e9 / peak}

    # ---- the timed result itself against the reference CPU filter: sampled chunks, downloaded after the timed region ----
    parity = None
    if rank == 0 and args.parity_chunks > 0:
        picks = sorted({0, n - 1, n // 3, (2 * n) // 3})[:args.parity_chunks]
        oks, kind = [], None
        for c in picks:
            ok, kind = check_against_cpu_filter(wl, maps[c].cpu().numpy(), download_chunk(buf, c, w, h))
            oks.append(ok)
        parity = {"chunks": [first_chunk + c for c in picks], "ok": all(oks), "against": kind,
                  "compared": "items, offsets, weight bits of every pixel of the sampled chunks"}
        if not all(oks):
            raise SystemExit(f"bench.py: the timed result differs from the CPU filter on chunks {picks}: {oks}")

    # ---- device-side consumer on the resident result (BASELINE.json config 4 chain) ----
    consumer = None
    if not args.no_consumer:
        gen = heightfield_generator(pkg, api, filt, wl.biomes)
        nc = min(n, 32)   # heights of at most 32 chunks per consumer launch measured (the kernel is compute-bound)
        heights = torch.empty((nc, h, w), dtype=torch.float32, device="cuda")
        offs = np.zeros((nc, 2), dtype=np.float32)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c_steps = 3
        for i in range(1 + c_steps):
            if i == 1:
                h0.record()
            gen(buf, 0, nc, offs, heights.data_ptr(), stream)
        h1.record()
        torch.cuda.synchronize()
        hms = h0.elapsed_time(h1) / c_steps * (n / nc)
        consumer = {"kernel": "heightfield_kernel", "ms": hms, "mpixels_per_s": n * w * h / (hms * 1e-3) / 1e6,
                    "octaves": 8, "bins_per_pixel": n_bins / (n * w * h), "chunks_measured": nc,
                    "chain_ms": ms_local + hms, "chain_mpixels_per_s": n * w * h / ((ms_local + hms) * 1e-3) / 1e6}
        gen.close()
        del heights
    c4 = None
    if rank == 0 and world == 1 and not args.no_consumer:
        c4 = config4_chain(pkg, api, workloads, filt, local, stream)

    # ---- end to end: host buffers in, page-locked host histograms out ----
    e2e = None
    if not args.no_e2e:
        sub = max(1, min(args.e2e_sub, n))
        host = torch.empty((n, th, tw), dtype=torch.uint16).pin_memory()
        host.copy_(maps)
        torch.cuda.synchronize()
        base_ptr = host.data_ptr()
        lib = api.library()
        U2 = ctypes.c_uint32 * 2
        groups = [(i, min(sub, n - i)) for i in range(0, n, sub)]
        n_threads = 2 if len(groups) > 1 else 1
        bufs = [FB(FB.STPExecutionType.Parallel) for _ in range(n_threads)]
        moved = [[0, 0, 0] for _ in range(n_threads)]
        errors = []

        def worker(t):
            try:
                api.stats_reset()
                for gi in range(t, len(groups), n_threads):
                    first, cnt = groups[gi]
                    ptrs = (ctypes.c_void_p * cnt)(*[base_ptr + (first + j) * th * tw * 2 for j in range(cnt)])
                    st = lib.shf_run_batch(filt._h, ptrs, cnt, U2(*wl.map_size), U2(*wl.nn), U2(*wl.total), bufs[t]._h,
                                           wl.radius)
                    if st != 0:
                        raise RuntimeError(lib.shf_last_error().decode())
                    hist = bufs[t].readHistogram()  # the step's result, read on the host
                    if int(hist.HistogramStartOffset[w * h]) <= 0:
                        raise RuntimeError("empty result")
                a, b_, c = api.stats()
                moved[t] = [a, b_, c]
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)

        def e2e_step():
            ts = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errors:
                raise errors[0]

        e2e_step()  # warm-up: grows the page-locked buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        dt = max_over_ranks(dt / args.e2e_steps)
        # bytes of the whole job: every rank moves its own share
        if world > 1:
            tb = torch.tensor([sum(m[1] for m in moved), sum(m[2] for m in moved)], dtype=torch.float64,
                              device="cuda" if args.backend == "nccl" else "cpu")
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            h2d_total, d2h_total = int(tb[0].item()), int(tb[1].item())
        else:
            h2d_total, d2h_total = sum(m[1] for m in moved), sum(m[2] for m in moved)
        e2e = {"value": total_chunks * w * h / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "d2h_gb_per_s": d2h_total / dt / 1e9,
               "how": f"shf_run_batch on pinned host maps, {sub} chunks per call, {n_threads} host threads with one "
                      "filter buffer each; bins + offsets copied to page-locked host memory inside the timed region",
               "gpu_launches_per_step": sum(m[0] for m in moved)}
        for b_ in bufs:
            b_.close()
        del host
        # the same box, the same moment: what plain concurrent D2H copies into page-locked memory reach
        ceil_all, ceil_rank = d2h_ceiling(barrier, sum_over_ranks)
        e2e["d2h_ceiling_gb_per_s"] = ceil_all
        e2e["d2h_ceiling_gb_per_s_this_rank_alone_share"] = ceil_rank
        e2e["frac_of_d2h_ceiling"] = e2e["d2h_gb_per_s"] / ceil_all

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        host_maps = [maps[i].cpu().numpy() for i in range(min(4, n))]
        cpu = cpu_filter_throughput(wl, host_maps, args.cpu_seconds, all_cores=False)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u16 samples, u32 counts, f32 weights", "data": "synthetic",
            "config": config_of(wl_named, world, args.scaling),
            "plan": plan, "bins_per_pixel": n_bins / (n * w * h),
            "ms_per_step_with_phase_events": ms_profiled,
            "step_call": "shf_run_device (host-blocking)" if args.sync_steps else
                         "shf_run_device_async back to back + shf_buffer_wait inside the timed region",
            "calls_repeated_on_checked_path": repeated,
            "roofline": roofline, "phases_ms": phases_ms, "parity_checked": parity, "cpu_baseline": cpu, "e2e": e2e,
            "consumer": consumer, "config4_chain": c4, "host_numa_node_rank0": numa,
            "gpu_launches": launches, "clocks": clocks, "impl": "ours",
        }
        emit(line)
    buf.close()
    filt.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: anything a library prints (e.g. NCCL's version banner) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
