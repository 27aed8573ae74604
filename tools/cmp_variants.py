#!/usr/bin/env python
"""Check that two builds of the library give bit-identical events: hashes of the band words and of the S1
columns for 2 x 3e6 events.  usage (on the GPU box): python tools/cmp_variants.py [a b]  with
nest_b200/csrc/variants/lib_<a>.so and lib_<b>.so built beforehand (e.g. -DNB_PE_NO_BOUNDS vs default:
the table bounds of pe_accept_exact must never change a decision)."""
import os, sys, subprocess, json
sys.path.insert(0, os.getcwd())
if len(sys.argv) == 2 and sys.argv[1] == "x":
    import numpy as np
    from nest_b200 import capi
    ctx = capi.Context(0)
    det = capi.detector("LUX_Run03"); rc = capi.prepare(det, 180.0); mo = capi.model(1); ana = capi.analysis(0)
    out = {}
    import hashlib
    for name, src in [("NR", capi.source(capi.SRC_FLAT, capi.NR, 0., 100., 180.)),
                      ("lowE", capi.source(capi.SRC_FLAT, capi.NR, 0., 6., 180.)),
                      ("CH3T", capi.source(capi.SRC_CH3T, capi.CH3T, 0., 18.5898, 180.)),
                      ("DD", capi.source(capi.SRC_DD, capi.DD, 0., 80., 180., par=(13., 0.12, 71.2, 20., -20.5))),
                      ("B8", capi.source(capi.SRC_B8, capi.B8, 0., 5., 180.)),
                      ("C14", capi.source(capi.SRC_C14, capi.C14, 0., 156., 180.)),
                      ("atmNu", capi.source(capi.SRC_ATMNU, capi.atmNu, 0., 85., 180.))]:
        band = capi.Band(ana)
        mo.massNum = det.molarMass if src.species >= 6 else 0.0
        keys = ("energy_kev", "z_mm", "s1_2", "s1_6", "s1_1", "s2_7")
        cols = ctx.run(det, mo, ana, src, rc, 5, 3_000_000, columns=list(keys), band=band)
        out[name] = [hashlib.sha256(band.words().tobytes()).hexdigest()[:16]] + [
            hashlib.sha256(cols[k].tobytes()).hexdigest()[:16] for k in keys]
    print(json.dumps(out))
else:
    res = {}
    names = tuple(sys.argv[1:3]) if len(sys.argv) >= 3 else ("nob", "bnd")
    for v in names:
        env = dict(os.environ, NB200_LIB=os.path.join(os.getcwd(), "nest_b200/csrc/variants/lib_%s.so" % v))
        res[v] = subprocess.run([sys.executable, __file__, "x"], env=env, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    a, b = json.loads(res[names[0]]), json.loads(res[names[1]])
    for k in a:
        print(k, "identical" if a[k] == b[k] else "DIFFERENT %r %r" % (a[k], b[k]))
