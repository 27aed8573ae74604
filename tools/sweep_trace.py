#!/usr/bin/env python
"""Where a config-4 sweep call spends its time: NB200_TRACE phases of nb200_run_sweep (host parameter blocks, launches,
kernels + band copy-back) for the bench's 250-point grid, next to one large run of the same source."""
import ctypes as C
import os
import sys
import time

os.environ["NB200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nest_b200 import capi  # noqa: E402

ctx = capi.Context(0)
ana = capi.analysis(0)
fields = [60., 100., 140., 180., 240., 300., 400., 500., 700., 1000.]
g1s = [0.09, 0.10, 0.117, 0.13, 0.15]
g1gs = [0.07, 0.085, 0.10, 0.115, 0.13]
wt = capi.WimpTable(50.0, 5.0, 0.0, seed=1)
grid = []
for f_ in fields:
    for g1_ in g1s:
        for gg_ in g1gs:
            d_ = capi.detector("LUX_Run03")
            d_.g1, d_.g1_gas = g1_, gg_
            grid.append((d_, capi.prepare(d_, f_), f_))
kp = len(grid)
npt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
for sname in ("WIMP50", "B8"):
    dets_ = [g[0] for g in grid]
    rcs_ = [g[1] for g in grid]
    mods_ = [capi.model(1) for _ in grid]
    if sname == "WIMP50":
        srcs_ = [wt.attach(capi.source(capi.SRC_WIMP, capi.WIMP, 50.0, 1e-36, g[2])) for g in grid]
    else:
        srcs_ = [capi.source(capi.SRC_B8, capi.B8, 0.0, 5.0, g[2]) for g in grid]
    Dk, Mk = (capi.Detector * kp)(*dets_), (capi.Model * kp)(*mods_)
    Sk, Rk = (capi.Source * kp)(*srcs_), (capi.RunConsts * kp)(*rcs_)
    hb = [capi.Band(ana) for _ in range(kp)]
    Bk = (capi.BandHist * kp)(*[b_.c for b_ in hb])
    for rep in range(3):
        t0 = time.perf_counter()
        ctx._check(capi.lib().nb200_run_sweep(ctx.h, kp, Dk, Mk, C.byref(ana), Sk, Rk, 1, rep * npt, npt, Bk))
        dt = time.perf_counter() - t0
        print("%s sweep rep %d: %.3f ms -> %.3e events/s" % (sname, rep, dt * 1e3, kp * npt / dt), file=sys.stderr)
    ctx.band_reset(ana)
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.run_async(dets_[0], mods_[0], ana, srcs_[0], rcs_[0], 1, 0, kp * npt)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        print("%s one run of %d events rep %d: %.3f ms -> %.3e events/s" % (sname, kp * npt, rep, dt * 1e3, kp * npt / dt), file=sys.stderr)
ctx.close()
