"""The public model methods of NESTcalc and the TestSpectra class in the host mirror (NEST.hh:256-333, 345, 404;
TestSpectra.hh:50-69) against the compiled reference: tests/cpp/class_surface.cpp calls them as reference-style code
does; deterministic results to 1e-12, stochastic ones by KS.  And the reference's own examples/bareNEST.cpp, compiled
UNMODIFIED against the mirror (oracle/_ref/bareNEST_b200, built where the reference tree is present), must run."""
import collections
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import RHO_LUX
from nest_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def surface(tmp_path_factory):
    host = os.path.join(ROOT, "nest_b200", "host")
    csrc = os.path.join(ROOT, "nest_b200", "csrc")
    r = subprocess.run(["make", "-s", "-C", host], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    exe = str(tmp_path_factory.mktemp("cs") / "class_surface")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + host,
           os.path.join(ROOT, "tests", "cpp", "class_surface.cpp"), "-L" + host, "-lnest_b200_host", "-L" + csrc, "-lnest_b200",
           "-Wl,-rpath," + host, "-Wl,-rpath," + csrc, "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-600:]
    rec = collections.defaultdict(list)
    for ln in r.stdout.splitlines():
        f = ln.split()
        rec[f[0]].append([float(x) for x in f[1:]])
    return {k: np.array(v) for k, v in rec.items()}


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_yield_methods_match_reference(surface, ref):
    mo = capi.model(0)
    nr, er = list(mo.NRYieldsParam), list(mo.ERYieldsParam)
    E = np.repeat([0.5, 3.0, 12.0, 41.5, 250.0], 3)
    F = np.tile([50., 180., 1000.], 5)
    # which: 0 dispatch, 1 BetaGR(multFact), 2 Gamma(multFact), 3 ERWeighted   (oracle/ref_probe.cpp ref_yields)
    cases = [("gamma", 2, capi.gammaRay, E, 1.0, 0., 0.), ("gamma1.2", 2, capi.gammaRay, E, 1.2, 0., 0.),
             ("erw", 3, capi.beta, E, 1.0, 0., 0.), ("nr", 0, capi.NR, E, 1.0, 131.293, 0.),
             ("ion", 0, capi.ion, E * 20., 1.0, 4., 2.), ("beta", 1, capi.beta, E, 1.0, 0., 0.),
             ("beta0.9", 1, capi.beta, E, 0.9, 0., 0.)]
    for tag, which, sp, en, mf, A, Z in cases:
        exp = ref.yields(which, sp, en, F, RHO_LUX, A, Z, nr, er, mf)
        got = surface[tag]
        assert got.shape == exp.shape == (15, 6)
        for k in range(5):
            nz = exp[:, k] != 0
            assert np.array_equal(got[:, k] == 0, exp[:, k] == 0), (tag, k)
            assert rel(got[nz, k], exp[nz, k]) < 1e-12, (tag, k)
    expv = ref.yields(0, capi.NR, [0.5, 3.0, 12.0, 41.5, 250.0], 180., RHO_LUX, 131.293, 0., nr, er, 1.0)
    assert rel(surface["nrvec"][:, :4], expv[:, :4]) < 1e-12
    # YieldResultValidity (NEST.cpp:1254-1271) on {-3, 1e9, ., 1.7}: Nph -> 0, Ne -> E / W_SCINT, L -> 1
    v = surface["valid"][0]
    assert v[0] == 0.0 and v[1] == 10.0 / 8.5e-3 and v[3] == 1.0


def test_width_model_matches_reference(ctx, surface, ref):
    # SURVEY 8c values of the reference: OmegaNR(0.3), OmegaER(180, 0.4, 372) default / CLI widths, FanoER(372, 180, CLI)
    assert rel(surface["omega"][0][:2], np.array([0.022985482790218961, 0.04626688043331114])) < 1e-12
    assert rel(surface["omegacli"][0][1:], np.array([0.045560206537141996, 0.16174786183636691])) < 1e-12
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    rng = np.random.default_rng(3)
    n = 5000
    ef, fr, nq = rng.uniform(1., 2000., n), rng.uniform(0.01, 0.99, n), np.exp(rng.uniform(np.log(2.), np.log(2e5), n))
    dp = C.POINTER(C.c_double)
    for which in (0, 1):
        mo = capi.model(which)
        w = np.array(list(mo.NRERWidthsParam)[:mo.n_widths])
        exp = np.zeros((n, 3))
        ref.L.ref_widths(C.c_int(n), ef.ctypes.data_as(dp), fr.ctypes.data_as(dp), nq.ctypes.data_as(dp), C.c_double(RHO_LUX),
                         w.ctypes.data_as(dp), C.c_int(w.size), exp.ctypes.data_as(dp))
        got = ctx.widths(det, mo, ef, fr, nq, RHO_LUX).T
        for k in range(3):
            assert rel(got[:, k], exp[:, k]) < 1e-12, (which, k)


def test_spike_and_xy_resolution_match_reference(ctx, surface, ref):
    dp = C.POINTER(C.c_double)
    # GetSpike: the reference on the same scintillation vector (needs a GetS1-shaped call path: ref_s1 covers the
    # in-chain use; here the law is zero-truncated N(scint[6], (sPEres/4) sqrt(scint[6])) and a fixed ratio)
    s = surface["spike"]
    mu, sg = 38.0, (0.37 / 4.) * np.sqrt(38.0)
    assert stats.kstest(s[:, 0], stats.norm(mu, sg).cdf).pvalue > 1e-4
    ratio = s[:, 1] / s[:, 0]
    s1 = ref.maps([50.0, 0.0], [-30.0, 0.0], [200.0, 544.95 - 1.5 * 160.0])[0]
    assert rel(ratio, np.full(len(ratio), s1[1] / s1[0])) < 1e-12
    # xyResolution against the reference's own draws
    n = 4000
    exp = np.zeros((n, 2))
    ref.L.ref_xyres(C.c_uint64(8), C.c_int(n), C.c_double(60.), C.c_double(-20.), C.c_double(900.), exp.ctypes.data_as(dp))
    got = surface["xy"]
    for k in range(2):
        assert stats.ks_2samp(got[:, k], exp[:, k]).pvalue > 1e-4
    r_g, r_r = np.hypot(got[:, 0] - 60., got[:, 1] + 20.), np.hypot(exp[:, 0] - 60., exp[:, 1] + 20.)
    assert stats.ks_2samp(r_g, r_r).pvalue > 1e-4
    det = capi.detector("LUX_Run03")
    a = ctx.xy_resolution(det, np.full(100, 60.), -20., 900., seed=3)
    b = ctx.xy_resolution(det, np.full(60, 60.), -20., 900., seed=3, first_event=40)
    assert np.array_equal(a[:, 40:], b)      # record i = event first_event + i


def test_testspectra_class_matches_reference(ctx, surface, ref):
    n = 3000
    for tag, kind, e0, e1 in (("ch3t", capi.SRC_CH3T, 0., 18.6), ("dd", capi.SRC_DD, 0., 80.), ("b8", capi.SRC_B8, 0., 5.),
                              ("ambe", capi.SRC_AMBE, 0.1, 200.), ("c14", capi.SRC_C14, 0., 156.)):
        exp = ref.spectrum(kind, 17, 20000, e0, e1)
        got = surface[tag][:, 0]
        assert got.size == n and got.min() >= 0
        assert stats.ks_2samp(got, exp).pvalue > 1e-4, tag
    # WIMP: the mirror's prep under its own isotope draws; energies against the reference sampler on the reference table
    integ, xmax, div = surface["wimpprep"][0]
    assert abs(integ / 1.0132e6 - 1) < 5e-3 and xmax == 90.0 and div == 0.2
    exp = ref.wimp_spectrum(5, 6, 20000, 50.0)
    assert stats.ks_2samp(surface["wimp"][:, 0], exp).pvalue > 1e-4
    assert 1e2 < surface["drate"][0][0] < 1e6
    # nb200_spectrum: draw i is the energy of event i of a chain run with the same source and seed
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    src = capi.source(capi.SRC_CH3T, capi.CH3T, 0.0, 18.5898, 180.0)
    mo = capi.model(1)
    mo.massNum = det.molarMass
    e_chain = ctx.run(det, mo, capi.analysis(0), src, rc, 9, 5000, first_event=100, columns=["energy_kev"])["energy_kev"]
    assert np.array_equal(ctx.spectrum(src, 5000, 9, first_event=100), e_chain)


def test_unmodified_bareNEST_runs_against_the_mirror():
    exe = os.path.join(ROOT, "oracle", "_ref", "bareNEST_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bareNEST_b200 not built (needs the reference tree at build time)")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-600:]
    rows = [ln.split("\t") for ln in r.stdout.splitlines() if ln[:1].isdigit()]
    assert len(rows) == 1 and len(rows[0]) == 12          # one event, execNEST's twelve columns
    keV, field, dt = float(rows[0][0]), float(rows[0][1]), float(rows[0][2])
    assert 0 <= keV < 200 and 50 < field < 500 and 38 <= dt <= 305   # NR 10-100 keV, FitEF field, LUX drift window
    assert "g1 = 0.117 phd per photon" in r.stdout                   # CalculateG2's banner
