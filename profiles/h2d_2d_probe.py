#!/usr/bin/env python
"""Strided (2-D) host-to-device copies of the last W bytes of every row of a pinned [n][stride] byte plane: rows per second and the
equivalent link bandwidth, against a plain copy of the same plane. usage: python profiles/h2d_2d_probe.py"""
import ctypes as C
import json
import time

rt = C.CDLL("libcudart.so.12")
def ck(e):
    assert e == 0, e
n, stride = 4_000_000, 150
h = C.c_void_p(); d = C.c_void_p(); s = C.c_void_p()
ck(rt.cudaSetDevice(0))
ck(rt.cudaHostAlloc(C.byref(h), C.c_size_t(n * stride), C.c_uint(1)))
C.memset(h, 1, n * stride)
ck(rt.cudaMalloc(C.byref(d), C.c_size_t(n * stride)))
ck(rt.cudaStreamCreate(C.byref(s)))
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
out = []
def timed(f, reps=5):
    f(); ck(rt.cudaStreamSynchronize(s))
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ck(rt.cudaStreamSynchronize(s))
    return (time.perf_counter() - t0) / reps
t = timed(lambda: ck(rt.cudaMemcpyAsync(d, h, C.c_size_t(n * stride), 1, s)))
out.append({"copy": "whole plane", "ms": round(t * 1e3, 2), "gbs": round(n * stride / t / 1e9, 1)})
for w in (16, 32, 48, 64):
    src = C.c_void_p(h.value + stride - w)
    t = timed(lambda: ck(rt.cudaMemcpy2DAsync(d, w, src, stride, w, n, 1, s)))
    out.append({"copy": f"last {w} bytes of every row (2-D)", "ms": round(t * 1e3, 2), "mrows_per_s": round(n / t / 1e6, 1), "payload_gbs": round(n * w / t / 1e9, 1)})
print(json.dumps(out))
