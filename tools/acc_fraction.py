#!/usr/bin/env python
"""In-window (band-accepted) fraction of config 1 / config 2 at very high GPU statistics, for any build of the
library (raw ctypes: works with older builds that lack newer exports).  usage: acc_fraction.py <lib.so> <events>"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nest_b200 import capi  # noqa: E402  (struct layouts only)

L = C.CDLL(sys.argv[1])
n = int(float(sys.argv[2]))
vp = C.c_void_p
h = vp()
L.nb200_create.argtypes = [C.c_int, C.POINTER(vp)]
assert L.nb200_create(0, C.byref(h)) == 0
det, rc, ana = capi.Detector(), capi.RunConsts(), capi.Analysis()
L.nb200_detector_builtin(b"LUX_Run03", C.byref(det))
L.nb200_prepare.argtypes = [C.POINTER(capi.Detector), C.c_double, C.c_double, C.POINTER(capi.RunConsts)]
assert L.nb200_prepare(C.byref(det), 180.0, 0.0, C.byref(rc)) == 0
L.nb200_analysis_default(0, C.byref(ana))
L.nb200_run_async.argtypes = [vp, C.POINTER(capi.Detector), C.POINTER(capi.Model), C.POINTER(capi.Analysis),
                              C.POINTER(capi.Source), C.POINTER(capi.RunConsts), C.c_uint64, C.c_uint64, C.c_uint64, vp]
L.nb200_band_reset.argtypes = [vp, C.POINTER(capi.Analysis)]
L.nb200_band_fetch.argtypes = [vp, C.POINTER(capi.BandHist)]
for name, sp, e0, e1, erm in (("ER_0.1-20", capi.beta, 0.1, 20.0, 0), ("NR_0-100", capi.NR, 0.0, 100.0, -1)):
    mo = capi.Model()
    L.nb200_model_default(1, C.byref(mo))
    mo.er_model = erm
    if sp >= 6:
        mo.massNum = det.molarMass
    src = capi.source(capi.SRC_FLAT, sp, e0, e1, 180.0)
    acc = tot = 0
    for seed in (11, 12, 13, 14):
        assert L.nb200_band_reset(h, C.byref(ana)) == 0
        done = 0
        while done < n // 4:
            m = min(100_000_000, n // 4 - done)
            assert L.nb200_run_async(h, C.byref(det), C.byref(mo), C.byref(ana), C.byref(src), C.byref(rc), seed, done, m, None) == 0
            done += m
        b = capi.Band(ana)
        assert L.nb200_band_fetch(h, C.byref(b.c)) == 0
        acc += int(b.accepted.sum())
        tot += done
    f = acc / tot
    print("%s %s: accepted %d of %d = %.6f +- %.6f" % (os.path.basename(sys.argv[1]), name, acc, tot, f, np.sqrt(f * (1 - f) / tot)))
