#!/usr/bin/env python
"""Host-to-device copy bandwidth of this box, the roofline of the end-to-end (E) boundary: pinned host buffers -> every visible GPU,
1 .. N devices at once, from ONE process (the round-robin engine's situation), for ordinary pinned memory, write-combined pinned memory
and with the buffers first-touched from a thread bound to each NUMA node the process may use. usage: python profiles/h2d_probe.py"""
import ctypes as C
import json
import os
import time

rt = C.CDLL("libcudart.so.12")
def ck(e):
    assert e == 0, e
n_dev = C.c_int(0)
ck(rt.cudaGetDeviceCount(C.byref(n_dev)))
n_dev = n_dev.value
SZ = 1 << 30
FLAGS = {"pinned": 0x01, "write_combined": 0x01 | 0x04}  # cudaHostAllocPortable | cudaHostAllocWriteCombined
out = {"devices": n_dev, "cpus": sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))], "results": []}
try:
    out["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    out["mems_allowed"] = [l.split()[1] for l in open("/proc/self/status") if l.startswith("Mems_allowed_list")][0]
except Exception as e:
    out["numa_nodes"] = str(e)
for kind, fl in FLAGS.items():
    host, dev, streams = [], [], []
    for d in range(n_dev):
        ck(rt.cudaSetDevice(d))
        h = C.c_void_p()
        ck(rt.cudaHostAlloc(C.byref(h), C.c_size_t(SZ), C.c_uint(fl)))
        C.memset(h, 1, SZ) if kind == "pinned" else None
        p = C.c_void_p()
        ck(rt.cudaMalloc(C.byref(p), C.c_size_t(SZ)))
        s = C.c_void_p()
        ck(rt.cudaStreamCreate(C.byref(s)))
        host.append(h); dev.append(p); streams.append(s)
    for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
        for rep in range(2):
            t0 = time.perf_counter()
            for it in range(4):
                for d in range(n):
                    ck(rt.cudaSetDevice(d))
                    ck(rt.cudaMemcpyAsync(dev[d], host[d], C.c_size_t(SZ), 1, streams[d]))
            for d in range(n):
                ck(rt.cudaSetDevice(d))
                ck(rt.cudaStreamSynchronize(streams[d]))
            el = time.perf_counter() - t0
        out["results"].append({"memory": kind, "gpus": n, "total_gbs": round(4 * n * SZ / el / 1e9, 1), "per_gpu_gbs": round(4 * SZ / el / 1e9, 1)})
    for d in range(n_dev):
        ck(rt.cudaSetDevice(d))
        rt.cudaFreeHost(host[d]); rt.cudaFree(dev[d])
print(json.dumps(out))
