"""Physics-level agreement the north star asks for: ER leakage below the NR band centroid per S1
bin (definition: reference examples/rootNEST.cpp:880-915), an AmBe power-law source, and the
loopNEST-style parameter sweep of BASELINE config 4 (grid of drift fields and g1 / g1_gas values,
one band per grid point)."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import gpu_chain, ref_chain
from nest_b200 import capi

pytestmark = pytest.mark.gpu


def leakage_per_bin(er, nr, ana):
    """fraction of ER events below the NR band mean, per S1 bin with enough statistics"""
    out = {}
    w = (ana.maxS1 - ana.minS1) / ana.numBins

    def sel(a):
        s1, s2 = a[:, refprobe.SIG1], a[:, refprobe.SIG2]
        ok = (s1 > 0) & (s2 > 0)
        return s1[ok], np.log10(s2[ok] / s1[ok])
    e1, ey = sel(er)
    n1, ny = sel(nr)
    for j in range(0, 40, 4):
        lo, hi = ana.minS1 + j * w, ana.minS1 + (j + 4) * w
        me, mn = (e1 > lo) & (e1 <= hi), (n1 > lo) & (n1 <= hi)
        if me.sum() < 2000 or mn.sum() < 500:
            continue
        centroid = ny[mn].mean()
        k = int((ey[me] < centroid).sum())
        out[j] = (k, int(me.sum()))
    return out


def test_er_leakage_below_nr_band(ctx, ref):
    n = 250000
    ana = capi.analysis(0)
    g_er, _ = gpu_chain(ctx, 41, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0)
    g_nr, _ = gpu_chain(ctx, 42, n, capi.NR, capi.SRC_FLAT, 0.0, 60.0)
    r_er = ref_chain(ref, 43, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0)
    r_nr = ref_chain(ref, 44, n, capi.NR, capi.SRC_FLAT, 0.0, 60.0)
    lg, lr = leakage_per_bin(g_er, g_nr, ana), leakage_per_bin(r_er, r_nr, ana)
    common = sorted(set(lg) & set(lr))
    assert len(common) >= 5
    tot_g = sum(lg[j][0] for j in common), sum(lg[j][1] for j in common)
    tot_r = sum(lr[j][0] for j in common), sum(lr[j][1] for j in common)
    for j in common:
        (kg, ng), (kr, nr_) = lg[j], lr[j]
        # two binomial proportions; the centroid itself carries statistical error, hence 5 sigma
        p = (kg + kr) / (ng + nr_)
        se = np.sqrt(max(p * (1 - p), 1e-9) * (1 / ng + 1 / nr_))
        assert abs(kg / ng - kr / nr_) < 5 * se + 2e-4, (j, kg / ng, kr / nr_, se)
    assert 0 < tot_g[0] and abs(tot_g[0] / tot_g[1] - tot_r[0] / tot_r[1]) < 0.25 * tot_r[0] / tot_r[1] + 5e-4


def test_ambe_power_law_source(ctx, ref):
    n = 100000
    g, _ = gpu_chain(ctx, 5, n, capi.AmBe, capi.SRC_AMBE, 0.5, 300.0, par=(-0.95351,))
    r = ref.spectrum  # the reference probe exposes CH3T/DD/B8 only: compare with the analytic law
    E = g[:, refprobe.COL["keV_true"]]
    assert E.min() >= 0.5 and E.max() <= 300.0
    # PowLawFit_spectrum (TestSpectra.cpp:151-167): x = (ymin + (ymax-ymin) u)^(1/(1+expo))
    q = 1 - 0.95351
    cdf = lambda x: (x ** q - 0.5 ** q) / (300.0 ** q - 0.5 ** q)
    assert stats.kstest(E, cdf).pvalue > 1e-4


def test_parameter_sweep_grid(ctx):
    """config 4 shape: WIMP 50 GeV and 8B over a grid of fields x g1 x g1_gas; every grid point is an
    independent run with its own flattened detector, per-run constants and band"""
    lib = capi.lib()
    dp = C.POINTER(C.c_double)
    base, expo = np.zeros(100), np.zeros(100)
    integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
    assert lib.nb200_wimp_prep(50.0, 5.0, 0.0, 7, base.ctypes.data_as(dp), expo.ctypes.data_as(dp), C.byref(integ),
                               C.byref(xmax), C.byref(div)) == 0
    ana, mo = capi.analysis(0), capi.model(1)
    n = 60000
    res = {}
    for field in (100.0, 400.0):
        for g1 in (0.10, 0.13):
            for g1_gas in (0.08, 0.11):
                det = capi.detector("LUX_Run03")
                det.g1, det.g1_gas = g1, g1_gas
                rc = capi.prepare(det, field)          # CalculateG2 per grid point
                assert abs(rc.g2_params[3] - rc.g2_params[0] * g1_gas * rc.g2_params[1]) < 1e-12
                for name, src in (("wimp", capi.source(capi.SRC_WIMP, capi.WIMP, 50.0, 1e-36, field)),
                                  ("b8", capi.source(capi.SRC_B8, capi.B8, 0.0, 5.0, field))):
                    if name == "wimp":
                        src.wimp_base, src.wimp_exponent = base.ctypes.data_as(dp), expo.ctypes.data_as(dp)
                        src.wimp_xMax, src.wimp_divisor, src.wimp_mass, src.wimp_day = xmax.value, div.value, 50.0, 0.0
                    band = capi.Band(ana)
                    cols = ctx.run(det, mo, ana, src, rc, 3, n, columns=["s1_5", "s2_7", "photons", "electrons"],
                                   band=band)
                    res[(name, field, g1, g1_gas)] = (np.abs(cols["s1_5"]).mean(), np.abs(cols["s2_7"]).mean(),
                                                     cols["photons"].mean(), cols["electrons"].mean(),
                                                     int(band.accepted.sum()))
    for name in ("wimp", "b8"):
        for field in (100.0, 400.0):
            # quanta do not depend on the optical gains (same seed: identical draws)
            q = {res[(name, field, a, b)][2:4] for a in (0.10, 0.13) for b in (0.08, 0.11)}
            assert len(q) == 1
            # S1 scales with g1, S2 with g1_gas (corrected S2 ~ g2)
            for b in (0.08, 0.11):
                r1 = res[(name, field, 0.13, b)][0] / res[(name, field, 0.10, b)][0]
                assert 1.2 < r1 < 1.45, (name, field, r1)
            for a in (0.10, 0.13):
                r2 = res[(name, field, a, 0.11)][1] / res[(name, field, a, 0.08)][1]
                assert 1.25 < r2 < 1.5, (name, field, r2)
        # more field -> less recombination -> fewer photons (the electron column also carries the
        # below-cathode zeroing of execNEST.cpp:1107, whose rate depends on the drift speed)
        assert res[(name, 400.0, 0.10, 0.08)][2] < res[(name, 100.0, 0.10, 0.08)][2]
    assert res[("wimp", 100.0, 0.13, 0.11)][4] > res[("b8", 100.0, 0.13, 0.11)][4]   # 8B barely makes the S1 window
