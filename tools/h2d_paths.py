"""Host -> device copy of a pageable numpy array: the library's threaded bounce-buffer staging against cudaHostRegister +
direct DMA, and the staging at several thread counts (set MLB200_COPY_THREADS per process: it is read per copy).
Usage: python tools/h2d_paths.py [gigabytes]"""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_b200 import cabi

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 3.2
d = 16
n = int(gb * 1e9 / (8 * d))
x = np.random.default_rng(0).random((n, d))
out = {"bytes": x.nbytes, "host_cores": os.cpu_count()}
ctx = cabi.Context(1)
for threads in (1, 2, 4, 8):
    os.environ["MLB200_COPY_THREADS"] = str(threads)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        data = cabi.Data.upload(ctx, x)
        ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
        data.close()
    out[f"staged_{threads}_threads_gbs"] = x.nbytes / best / 1e9
del os.environ["MLB200_COPY_THREADS"]

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
rt.cudaFree.argtypes = [ctypes.c_void_p]
dev = ctypes.c_void_p()
assert rt.cudaMalloc(ctypes.byref(dev), x.nbytes) == 0
for rep in range(2):
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(x.ctypes.data, x.nbytes, 0)
    t1 = time.perf_counter()
    assert rc == 0, rc
    assert rt.cudaMemcpy(dev, x.ctypes.data, x.nbytes, 1) == 0
    rt.cudaDeviceSynchronize()
    t2 = time.perf_counter()
    rt.cudaHostUnregister(x.ctypes.data)
    t3 = time.perf_counter()
    out[f"register_rep{rep}"] = {"register_s": t1 - t0, "copy_s": t2 - t1, "unregister_s": t3 - t2, "copy_gbs": x.nbytes / (t2 - t1) / 1e9,
                                 "total_gbs": x.nbytes / (t3 - t0) / 1e9}
t0 = time.perf_counter()
assert rt.cudaMemcpy(dev, x.ctypes.data, x.nbytes, 1) == 0
rt.cudaDeviceSynchronize()
out["plain_cudaMemcpy_pageable_gbs"] = x.nbytes / (time.perf_counter() - t0) / 1e9
rt.cudaFree(dev)
print(json.dumps(out))
