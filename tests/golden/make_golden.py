#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref/libnestprobe.so, i.e. /root/reference):
    python tests/golden/make_golden.py
The reference ships no golden vectors for this path; these are outputs of the
compiled reference itself (LUX_Run03 detector, seeds given below) and pin the oracle port
wherever the reference tree is absent (the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refprobe  # noqa: E402
from helpers import RHO_LUX, ref_chain  # noqa: E402
from nest_b200 import capi  # noqa: E402


def main():
    r = refprobe.reference()
    assert r is not None, "build oracle/_ref first (make -C oracle ref)"
    mo = capi.model(1)
    nr, er, w = list(mo.NRYieldsParam), list(mo.ERYieldsParam), list(mo.NRERWidthsParam)[:13]
    out = {}
    # deterministic yields: species x energy x field grid
    E = np.exp(np.linspace(np.log(0.02), np.log(500.0), 160))
    F = np.array([1.0, 50.0, 180.0, 400.0, 1000.0, 4000.0])
    EE, FF = [a.ravel().copy() for a in np.meshgrid(E, F, indexing="ij")]
    out["y_E"], out["y_F"] = EE, FF
    for name, which, sp, mult in [("NR", 0, capi.NR, 1.0), ("beta", 0, capi.beta, 1.0), ("betaGR09", 1, capi.beta, 0.9),
                                  ("gamma", 0, capi.gammaRay, 1.0), ("gamma13", 2, capi.beta, 1.3),
                                  ("weighted", 3, capi.beta, 1.0)]:
        out["y_" + name] = r.yields(which, sp, EE, FF, RHO_LUX, 131.293, 54.0, nr, er, mult)
    # Kr83m: the three lines x decay separations (min == max => no random draw) x fields
    Fk = np.array([1.0, 50.0, 180.0, 400.0, 1000.0, 4000.0])
    out["y_kr_F"] = Fk
    for e in (9.4, 32.1, 41.5):
        for dT in (50.0, 154.4, 400.0, 1000.0):
            out["y_kr_%g_%g" % (e, dT)] = r.yields(0, capi.Kr83m, np.full(Fk.size, e), Fk, RHO_LUX, dT, dT, nr, er)
    Ea = np.linspace(500.0, 9000.0, 60)
    out["y_alpha_E"] = Ea
    out["y_alpha"] = r.yields(0, capi.ion, Ea, 180.0, RHO_LUX, 4.0, 2.0, nr, er)
    # per-run constants: rho, Wq, alpha, vD(180), g2_params[5], NewMean
    import ctypes as C
    c = np.zeros(10)
    r.L.ref_consts(C.c_double(180.0), c.ctypes.data_as(C.c_void_p))
    out["consts_180"] = c
    x = np.linspace(-180, 180, 41)
    y = np.linspace(120, -150, 41)
    z = np.linspace(5, 540, 41)
    out["maps_xyz"] = np.stack([x, y, z])
    out["maps"] = np.stack(r.maps(x, y, z))
    # the sequential chain at fixed seeds (bit pins for the oracle port)
    cases = {
        "chain_ER_flat": (1, 400, capi.beta, capi.SRC_FLAT, 0.1, 20.0, 180.0, 0, (0, 0, 0), 0),
        "chain_NR_flat": (2, 400, capi.NR, capi.SRC_FLAT, 0.0, 100.0, 180.0, 0, (0, 0, 0), -1),
        "chain_CH3T": (3, 300, capi.CH3T, capi.SRC_CH3T, 0.0, 18.5898, 180.0, 0, (0, 0, 0), -1),
        "chain_DD": (4, 300, capi.DD, capi.SRC_DD, 0.0, 80.0, 180.0, 0, (0, 0, 0), -1),
        "chain_B8": (5, 300, capi.B8, capi.SRC_B8, 0.0, 5.0, 180.0, 0, (0, 0, 0), -1),
        "chain_ER_fitEF": (6, 200, capi.beta, capi.SRC_FLAT, 1.0, 15.0, -1.0, 0, (0, 0, 0), 0),
        "chain_gamma_mono": (7, 200, capi.gammaRay, capi.SRC_MONO, 41.5, 41.5, 180.0, 1, (10., 20., 300.), -1),
        "chain_NR_250_parametric": (8, 200, capi.NR, capi.SRC_MONO, 250.0, 250.0, 180.0, 1, (10., 20., 300.), -1),
        "chain_alpha": (9, 100, capi.ion, capi.SRC_MONO, 5300.0, 5300.0, 180.0, 1, (10., 20., 300.), -1),
        # Kr83m: eMin = the line, eMax = max time separation [ns] (execNEST.cpp:583-588)
        "chain_Kr83m_41": (10, 200, capi.Kr83m, capi.SRC_MONO, 41.5, 1000.0, 180.0, 0, (0, 0, 0), -1),
        "chain_Kr83m_9": (11, 200, capi.Kr83m, capi.SRC_MONO, 9.4, 400.0, 180.0, 0, (0, 0, 0), -1),
        "chain_Kr83m_32": (12, 200, capi.Kr83m, capi.SRC_MONO, 32.1, 400.0, 180.0, 0, (0, 0, 0), -1),
        "chain_C14": (13, 200, capi.C14, capi.SRC_C14, 0.0, 156.0, 180.0, 0, (0, 0, 0), -1),
        "chain_Cf": (14, 200, capi.Cf, capi.SRC_CF, 0.0, 200.0, 180.0, 0, (0, 0, 0), -1),
        "chain_ppSolar": (15, 200, capi.ppSolar, capi.SRC_PPSOLAR, 0.0, 250.0, 180.0, 0, (0, 0, 0), -1),
        "chain_atmNu": (16, 200, capi.atmNu, capi.SRC_ATMNU, 0.0, 85.0, 180.0, 0, (0, 0, 0), -1),
        "chain_AmBe": (17, 200, capi.AmBe, capi.SRC_AMBE, 0.5, 300.0, 180.0, 0, (0, 0, 0), -1),
    }
    for name, (seed, n, sp, kind, e0, e1, fld, fp, xyz, erm) in cases.items():
        out[name] = ref_chain(r, seed, n, sp, kind, e0, e1, fld, fp, xyz, erm)
    out["chain_cases"] = np.array(repr(cases))
    # band of a longer ER run
    big = ref_chain(r, 11, 20000, capi.beta, capi.SRC_FLAT, 0.1, 20.0, 180.0, 0, (0, 0, 0), 0)
    out["band_sig"] = np.stack([big[:, refprobe.SIG1], big[:, refprobe.SIG2]])
    out["band"] = r.getband(big[:, refprobe.SIG1], big[:, refprobe.SIG2])
    # samplers
    for kind, p in [(0, (0, 0, 0)), (1, (3., 2., 1.)), (2, (0.9962, 0.37, 0)), (3, (100., 10., 2.25)),
                    (4, (1000, 0.117, 0)), (4, (20, 0.3, 0)), (5, (154.4, 0., 1000.)), (6, (0, 0, 0)), (7, (55.5, 0, 0))]:
        out["sampler_%d_%g_%g" % (kind, p[0], p[1])] = r.sampler(kind, 5, 400, *p)
    for k, nme in [(capi.SRC_CH3T, "CH3T"), (capi.SRC_DD, "DD"), (capi.SRC_B8, "B8")]:
        out["spectrum_" + nme] = r.spectrum(k, 9, 400, 0.0, 80.0)
    for k, nme, hi in [(capi.SRC_C14, "C14", 156.0), (capi.SRC_CF, "Cf", 200.0), (capi.SRC_PPSOLAR, "ppSolar", 250.0),
                       (capi.SRC_ATMNU, "atmNu", 85.0), (capi.SRC_AMBE, "AmBe", 300.0)]:
        out["spectrum_" + nme] = r.spectrum(k, 9, 400, 0.5, hi)
    # LAr
    El = np.exp(np.linspace(np.log(0.1), np.log(1000.0), 300))
    out["lar_E"] = El
    for sp in (0, 1, 2):
        for fld in (1.0, 500.0, 1000.0):
            yv, fv = r.lar(4, sp, El, fld, 1.393)
            out["lar_y_%d_%g" % (sp, fld)] = yv
            out["lar_f_%d_%g" % (sp, fld)] = fv
    # runNESTvec
    xyz = np.column_stack([np.linspace(-100, 100, 200), np.linspace(50, -80, 200), np.linspace(60, 500, 200)])
    Ev = np.linspace(0.5, 90, 200)
    m0 = capi.model(0)
    out["vec_xyz"], out["vec_E"] = xyz, Ev
    out["vec_beta_180"] = r.runvec(capi.beta, Ev, xyz, 180.0, 3, list(m0.ERYieldsParam), list(m0.NRYieldsParam),
                                   list(m0.NRERWidthsParam)[:13])
    out["vec_NR_fitEF"] = r.runvec(capi.NR, Ev, xyz, -1.0, 3, list(m0.ERYieldsParam), list(m0.NRYieldsParam),
                                   list(m0.NRERWidthsParam)[:13])
    np.savez_compressed(os.path.join(HERE, "reference_lux_run03.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
