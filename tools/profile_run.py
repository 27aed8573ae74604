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
ap.add_argument("--config", default="NR", choices=["NR", "ER", "CH3T", "DD"])
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
