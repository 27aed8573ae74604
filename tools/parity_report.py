#!/usr/bin/env python
"""Measured parity numbers for profiles/: run on the GPU box (needs oracle/_ref prebuilt).
   python tools/parity_report.py > gpurun_out/parity_report.txt"""
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refprobe  # noqa: E402
from helpers import RHO_LUX, COLUMN_NAMES, band_chi2, gpu_chain, ks_report, ref_chain  # noqa: E402
from nest_b200 import capi  # noqa: E402

ctx = capi.Context(0)
ref = refprobe.reference()
det = capi.detector("LUX_Run03")
capi.prepare(det, 180.0)
print("# nest-b200 parity report (GPU = B200 through the C-ABI; reference = unmodified NEST, oracle/_ref)\n")
print("## deterministic mean yields: max relative error over 400 energies x 7 fields x 3 densities")
E = np.exp(np.linspace(np.log(0.01), np.log(3000.0), 400))
F = np.array([1.0, 10.0, 50.0, 180.0, 400.0, 1000.0, 4000.0])
EE, FF = [a.ravel().copy() for a in np.meshgrid(E, F, indexing="ij")]
for name, species, er_model, which_ref, mult in [("NR", capi.NR, -1, 0, 1.0), ("beta", capi.beta, -1, 0, 1.0),
                                                 ("betaGR x0.9", capi.beta, 0, 1, 0.9), ("gamma", capi.gammaRay, -1, 0, 1.0),
                                                 ("gamma x1.3", capi.beta, 1, 2, 1.3), ("ER weighted", capi.beta, 2, 3, 1.0)]:
    mo = capi.model(1)
    mo.er_model, mo.multFact, mo.massNum, mo.atomNum = er_model, mult, 131.293, 54.0
    worst = 0.0
    ee = EE if name != "NR" else np.minimum(EE, 330.0)
    for rho in (RHO_LUX, 2.7, 3.05):
        got = ctx.yields(det, mo, species, ee, FF, rho)
        exp = ref.yields(which_ref, species, ee, FF, rho, 131.293, 54.0, list(mo.NRYieldsParam), list(mo.ERYieldsParam), mult).T
        for k in range(4):
            d = np.abs(got[k] - exp[k]) / np.maximum(np.maximum(np.abs(got[k]), np.abs(exp[k])), 1e-300)
            worst = max(worst, d.max())
    print("  %-12s %.2e" % (name, worst))
print("\n## liquid argon: max relative error over the 50 000-energy benchmark grid, F = 1, 100, 500, 1000 V/cm")
Eg = .1 + (1000 - .1) / 50000 * np.arange(50000)
for sp, nm in enumerate(["NR", "ER", "Alpha"]):
    worst = 0.0
    for Fv in (1.0, 100.0, 500.0, 1000.0):
        y, _ = ctx.lar(sp, Eg, Fv, 1.393, want_fluct=False)
        ry, _ = ref.lar(1, sp, Eg, Fv, 1.393)
        d = np.abs(y.T - ry) / np.maximum(np.maximum(np.abs(y.T), np.abs(ry)), 1e-300)
        worst = max(worst, d.max())
    print("  %-6s %.2e" % (nm, worst))
y, _ = ctx.lar(1, Eg, 500.0, 1.393, want_fluct=False)
ry, _ = ref.lar(1, 1, Eg, 500.0, 1.393)
print("  ER `Lindhard == 1.` branch: GPU true for %d / 50000 energies, reference %d, identical mask: %s" %
      ((y[7] == 1).sum(), (ry[:, 7] == 1).sum(), np.array_equal(y[7] == 1, ry[:, 7] == 1)))
print("\n## stochastic chain, 2e5 events each side: smallest two-sample KS p-value over the 36 output columns")
for label, sp, kind, e0, e1, erm in [("config 1: ER flat 0.1-20 keV", capi.beta, capi.SRC_FLAT, 0.1, 20.0, 0),
                                     ("config 2: NR flat 0-100 keV", capi.NR, capi.SRC_FLAT, 0.0, 100.0, -1),
                                     ("CH3T", capi.CH3T, capi.SRC_CH3T, 0.0, 18.5898, -1),
                                     ("DD", capi.DD, capi.SRC_DD, 0.0, 80.0, -1)]:
    par = (13., 0.12, 71.2, 20., -20.5) if kind == capi.SRC_DD else ()
    ana = capi.analysis(0)
    band = capi.Band(ana)
    g, _ = gpu_chain(ctx, 3, 200000, sp, kind, e0, e1, er_model=erm, band=band, par=par)
    r = ref_chain(ref, 5, 200000, sp, kind, e0, e1, er_model=erm)
    ks = ks_report(g, r)
    worst = min(ks, key=ks.get)
    cm, cw, dof = band_chi2(band.finalize(), ref.getband(r[:, refprobe.SIG1], r[:, refprobe.SIG2]))
    print("  %-30s min KS p = %.3f (%s); band chi2/dof: mean %.2f, width %.2f (dof %d)" % (label, ks[worst], worst, cm, cw, dof))
ctx.close()
