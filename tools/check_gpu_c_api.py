"""Drives libmultiverso_gpu.so (the reference's float C API served by the device plane) through ctypes
exactly like the reference's Python binding drives libmultiverso.so; run on a GPU box, 1..8 ranks:

    python tools/check_gpu_c_api.py
    python tools/mvrun.py -n 2 -- python tools/check_gpu_c_api.py
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "multiverso_b200", "_lib", "libmultiverso_gpu.so"), mode=ctypes.RTLD_GLOBAL)
lib.MV_Init(None, None)
W = lib.MV_NumWorkers()
fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

h = ctypes.c_void_p()
lib.MV_NewArrayTable(1000, ctypes.byref(h))
d = np.arange(1000, dtype=np.float32)
lib.MV_AddArrayTable(h, fp(d), 1000)
lib.MV_AddAsyncArrayTable(h, fp(d), 1000)
lib.MV_Barrier()
out = np.zeros(1000, np.float32)
lib.MV_GetArrayTable(h, fp(out), 1000)
assert np.array_equal(out, d * 2 * W), out[:5]

m = ctypes.c_void_p()
lib.MV_NewMatrixTable(11, 8, ctypes.byref(m))
base = np.arange(88, dtype=np.float32).reshape(11, 8)
lib.MV_AddMatrixTableAll(m, fp(base), 88)
rows = (ctypes.c_int * 4)(0, 1, 5, 10)
sub = np.ascontiguousarray(base[[0, 1, 5, 10]])
lib.MV_AddMatrixTableByRows(m, fp(sub), 32, rows, 4)
lib.MV_Barrier()
full = np.zeros(88, np.float32)
lib.MV_GetMatrixTableAll(m, fp(full), 88)
expect = base * W
expect[[0, 1, 5, 10]] *= 2
assert np.array_equal(full.reshape(11, 8), expect)
got = np.zeros(32, np.float32)
lib.MV_GetMatrixTableByRows(m, fp(got), 32, rows, 4)
assert np.array_equal(got.reshape(4, 8), expect[[0, 1, 5, 10]])
lib.MV_Barrier()
print(f"gpu c api ok (worker {lib.MV_WorkerId()} of {W})", flush=True)
lib.MV_ShutDown()
