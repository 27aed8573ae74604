#!/usr/bin/env python
"""k_band_fit on the bands of an NR sweep (250 grid points x 1e5 events, 98 S1 bins x 30 log bins each): wall clock of the
C-ABI call (result copy-back included); run under `ncu --metrics gpu__time_duration.sum -k regex:k_band_fit` for the
kernel alone."""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from nest_b200 import capi  # noqa: E402

ctx = capi.Context(0)
ana = capi.analysis(0)
fields = [60., 100., 140., 180., 240., 300., 400., 500., 700., 1000.]
dets, mods, srcs, rcs = [], [], [], []
for f, g1, g1g in itertools.product(fields, [0.09, 0.10, 0.117, 0.13, 0.15], [0.07, 0.085, 0.10, 0.115, 0.13]):
    det = capi.detector("LUX_Run03")
    det.g1, det.g1_gas = g1, g1g
    rcs.append(capi.prepare(det, f))
    dets.append(det)
    mods.append(capi.model(1))
    srcs.append(capi.source(capi.SRC_FLAT, capi.NR, 0.0, 60.0, f))
bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=1, n=100_000)
k = len(bands)
line = np.full(ana.numBins, 2.0)
for model in (0, 1):
    ctx.sweep_band_fit(k, ana, model=model, line=line)
    t0 = time.perf_counter()
    for rep in range(5):
        fit, leak = ctx.sweep_band_fit(k, ana, model=model, line=line)
    dt = (time.perf_counter() - t0) / 5
    nfit = int((fit[:, :, -1] > 0).sum())
    print("model %d: %d slices, %d fitted, %.3f ms per call (wall, copy-back included)" % (model, k * ana.numBins, nfit, dt * 1e3))
t0 = time.perf_counter()
for b in bands[:10]:
    b.fit_gauss()
print("host nb200_band_fit_gauss: %.3f ms for %d bands (scaled from 10)" % ((time.perf_counter() - t0) * 1e3 * k / 10, k))
ctx.close()
