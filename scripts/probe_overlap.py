"""Probe: do two device-resident calls on two streams overlap (small latency-bound kernels of one under the emit of the other)?"""
import dataclasses, sys, threading, time
sys.path.insert(0, ".")
import torch
import superterrainplus_b200 as pkg
from superterrainplus_b200 import workloads

wl = dataclasses.replace(workloads.CONFIGS["C3"], dist=sys.argv[1] if len(sys.argv) > 1 else "uniform", chunks=128)
dev = torch.device("cuda", 0)
maps = [workloads.make_maps_torch(wl, i * 128, 128, dev) for i in range(2)]
info = pkg.STPNearestNeighbourInformation(wl.map_size, wl.nn, wl.total)
filt = pkg.STPSingleHistogramFilter(0)
FB = pkg.STPSingleHistogramFilter.STPFilterBuffer
bufs = [FB(FB.STPExecutionType.Parallel) for _ in range(2)]
lo, hi = torch.cuda.Stream.priority_range()
streams = [torch.cuda.Stream(priority=lo), torch.cuda.Stream(priority=hi)]
tw, th = wl.total

def run(i, reps):
    for _ in range(reps):
        filt.runDevice(maps[i].data_ptr(), th * tw, 128, info, bufs[i], wl.radius, streams[i].cuda_stream)

for i in range(2):
    run(i, 2)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(0, 10); run(1, 10); torch.cuda.synchronize(); seq = time.perf_counter() - t0
t0 = time.perf_counter()
ts = [threading.Thread(target=run, args=(i, 10)) for i in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]; torch.cuda.synchronize(); par = time.perf_counter() - t0
print(f"{wl.dist}: sequential {seq*1e3/20:.3f} ms per 128-chunk call, two streams concurrently {par*1e3/20:.3f} ms per call ({seq/par:.3f}x)")
