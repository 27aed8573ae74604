"""Helper of test_gpu_variants.py: hashes of the band words and of a few event columns for several
source configurations, computed with whichever library NB200_LIB selects (one library per process)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nest_b200 import capi  # noqa: E402

ctx = capi.Context(0)
det = capi.detector("LUX_Run03")
rc = capi.prepare(det, 180.0)
ana = capi.analysis(0)
out = {}
CASES = [("NR", capi.SRC_FLAT, capi.NR, 0., 100., ()), ("NR_lowE", capi.SRC_FLAT, capi.NR, 0., 6., ()),
         ("ER", capi.SRC_FLAT, capi.beta, 0.1, 20., ()), ("CH3T", capi.SRC_CH3T, capi.CH3T, 0., 18.5898, ()),
         ("DD", capi.SRC_DD, capi.DD, 0., 80., (13., 0.12, 71.2, 20., -20.5)), ("B8", capi.SRC_B8, capi.B8, 0., 5., ()),
         ("C14", capi.SRC_C14, capi.C14, 0., 156., ()), ("atmNu", capi.SRC_ATMNU, capi.atmNu, 0., 85., ())]
KEYS = ("energy_kev", "z_mm", "photons", "electrons", "s1_0", "s1_1", "s1_2", "s1_6", "s2_0", "s2_3", "s2_7", "signalE")
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402

wbase, wexpo = np.zeros(100), np.zeros(100)
integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
dp = C.POINTER(C.c_double)
assert capi.lib().nb200_wimp_prep(50.0, 5.0, 0.0, 11, wbase.ctypes.data_as(dp), wexpo.ctypes.data_as(dp), C.byref(integ),
                                  C.byref(xmax), C.byref(div)) == 0
CASES.append(("WIMP50", capi.SRC_WIMP, capi.WIMP, 0., 0., ()))
for name, kind, sp, e0, e1, par in CASES:
    mo = capi.model(1)
    mo.massNum = det.molarMass if sp >= 6 else 0.0
    if name == "ER":
        mo.er_model = 0
    src = capi.source(kind, sp, e0, e1, 180.0, par=par)
    if kind == capi.SRC_WIMP:
        src.wimp_base, src.wimp_exponent = wbase.ctypes.data_as(dp), wexpo.ctypes.data_as(dp)
        src.wimp_xMax, src.wimp_divisor, src.wimp_mass, src.wimp_day = xmax.value, div.value, 50.0, 0.0
    band = capi.Band(ana)
    cols = ctx.run(det, mo, ana, src, rc, 5, int(sys.argv[1]), columns=list(KEYS), band=band)
    out[name] = [hashlib.sha256(band.words().tobytes()).hexdigest()[:16]] + [
        hashlib.sha256(cols[k].tobytes()).hexdigest()[:16] for k in KEYS]
print(json.dumps(out))
