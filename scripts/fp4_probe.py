#!/usr/bin/env python
"""Is kind::mxf4 (e2m1 operands, f32 accumulate) usable for this integer contraction?  Prints the loads-free
pipe rate next to kind::i8 and whether long accumulations of +1/-1 products stay exact below 2^24."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccphylo_b200 import api  # noqa: E402

L = api.load()
L.ccg_measure_fp4_peak.restype = C.c_double
L.ccg_measure_fp4_peak.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
ctx = api.Context(0)
for check in (1e4, 1e6, 1.6e7, 3.0e7):
    bad = C.c_longlong(-1)
    info = (C.c_int * 3)()
    tops = L.ccg_measure_fp4_peak(ctx._h, 10.0, check, C.byref(bad), info)
    print(f"check_sum={check:.3g}: dot32={info[0]} iters={info[2]} D[0][0]={info[1]} expected={info[2] * 8 * info[0]} "
          f"inexact_elements={bad.value}  fp4 burst {tops:.0f} TOP/s", file=sys.stderr)
print(f"fp4 sustained (600 ms): {L.ccg_measure_fp4_peak(ctx._h, 600.0, 1e4, None, None):.0f} TOP/s", file=sys.stderr)
print(f"int8 burst {ctx.measure_i8_peak(10.0):.0f} TOP/s, sustained (600 ms) {ctx.measure_i8_peak(600.0):.0f} TOP/s", file=sys.stderr)
ctx.close()
