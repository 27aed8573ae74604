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
    for name, src in [("NR", capi.source(capi.SRC_FLAT, capi.NR, 0., 100., 180.)), ("lowE", capi.source(capi.SRC_FLAT, capi.NR, 0., 6., 180.))]:
        band = capi.Band(ana)
        cols = ctx.run(det, mo, ana, src, rc, 5, 3_000_000, columns=["s1_2", "s1_6", "s1_1"], band=band)
        import hashlib
        out[name] = [hashlib.sha256(band.words().tobytes()).hexdigest()] + [hashlib.sha256(cols[k].tobytes()).hexdigest() for k in ("s1_2", "s1_6", "s1_1")]
    print(json.dumps(out))
else:
    res = {}
    names = tuple(sys.argv[1:3]) if len(sys.argv) >= 3 else ("nob", "bnd")
    for v in names:
        env = dict(os.environ, NB200_LIB=os.path.join(os.getcwd(), "nest_b200/csrc/variants/lib_%s.so" % v))
        res[v] = subprocess.run([sys.executable, __file__, "x"], env=env, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    print(res[names[0]] == res[names[1]], res[names[1]][:200])
