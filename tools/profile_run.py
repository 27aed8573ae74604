#!/usr/bin/env python
"""Small driver for ncu: a few band-only launches of the chain kernel (BASELINE config 2 by default)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nest_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=float, default=4e6)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--config", default="NR", choices=["NR", "ER", "CH3T", "DD", "B8", "WIMP"])
a = ap.parse_args()
ctx = capi.Context(0)
det = capi.detector("LUX_Run03")
rc = capi.prepare(det, 180.0)
mo = capi.model(1)
ana = capi.analysis(0)
if a.config == "NR":
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
elif a.config == "ER":
    src = capi.source(capi.SRC_FLAT, capi.beta, 0.1, 20.0, 180.0)
    mo.er_model = 0
elif a.config == "CH3T":
    src = capi.source(capi.SRC_CH3T, capi.CH3T, 0.0, 18.5898, 180.0)
elif a.config == "B8":
    src = capi.source(capi.SRC_B8, capi.B8, 0.0, 5.0, 180.0)
elif a.config == "WIMP":  # 50 GeV WIMP, TestSpectra::WIMP_prep_spectrum(50, E_step 5, day 0)
    import ctypes as C
    import numpy as np
    base, expo = np.zeros(100), np.zeros(100)
    integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
    dp = C.POINTER(C.c_double)
    assert capi.lib().nb200_wimp_prep(50.0, 5.0, 0.0, 11, base.ctypes.data_as(dp), expo.ctypes.data_as(dp), C.byref(integ),
                                      C.byref(xmax), C.byref(div)) == 0
    src = capi.source(capi.SRC_WIMP, capi.WIMP, 0.0, 0.0, 180.0)
    src.wimp_base, src.wimp_exponent = base.ctypes.data_as(dp), expo.ctypes.data_as(dp)
    src.wimp_xMax, src.wimp_divisor, src.wimp_mass, src.wimp_day = xmax.value, div.value, 50.0, 0.0
else:
    src = capi.source(capi.SRC_DD, capi.DD, 0.0, 80.0, 180.0, par=(13., 0.12, 71.2, 20., -20.5))
ctx.band_reset(ana)
n = int(a.events)
for i in range(a.launches):
    t0 = time.perf_counter()
    ctx.run_async(det, mo, ana, src, rc, 1, i * n, n)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("launch %d: %d events in %.3f ms -> %.3e events/s" % (i, n, dt * 1e3, n / dt))
band = capi.Band(ana)
ctx.band_fetch(band)
print("accepted", int(band.accepted.sum()), "rejected", int(band.rejected.sum()))
ctx.close()
