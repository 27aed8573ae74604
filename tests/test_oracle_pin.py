"""The oracle is pinned before it is trusted: the CPU restatement (oracle/liboracle.so) must
reproduce the UNMODIFIED reference bit for bit -- against the committed golden fixtures
(tests/golden/reference_lux_run03.npz, outputs of the compiled reference) everywhere, and against
the reference library / binary themselves wherever oracle/_ref has been built."""
import ast
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refprobe
from helpers import RHO_LUX, ref_chain
from nest_b200 import capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_lux_run03.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_yields_bitwise(orc, gold):
    mo = capi.model(1)
    nr, er = list(mo.NRYieldsParam), list(mo.ERYieldsParam)
    E, F = gold["y_E"], gold["y_F"]
    for name, which, sp, mult in [("NR", 0, capi.NR, 1.0), ("beta", 0, capi.beta, 1.0), ("betaGR09", 1, capi.beta, 0.9),
                                  ("gamma", 0, capi.gammaRay, 1.0), ("gamma13", 2, capi.beta, 1.3),
                                  ("weighted", 3, capi.beta, 1.0)]:
        got = orc.yields(which, sp, E, F, RHO_LUX, 131.293, 54.0, nr, er, mult)
        assert np.array_equal(got, gold["y_" + name]), name
    got = orc.yields(0, capi.ion, gold["y_alpha_E"], 180.0, RHO_LUX, 4.0, 2.0, nr, er)
    assert np.array_equal(got, gold["y_alpha"])
    Fk = gold["y_kr_F"]
    for e in (9.4, 32.1, 41.5):  # GetYieldKr83m, fixed decay separation (NEST.cpp:954-1044)
        for dT in (50.0, 154.4, 400.0, 1000.0):
            got = orc.yields(0, capi.Kr83m, np.full(Fk.size, e), Fk, RHO_LUX, dT, dT, nr, er)
            assert np.array_equal(got, gold["y_kr_%g_%g" % (e, dT)]), (e, dT)


def test_run_constants_and_maps_bitwise(orc, gold):
    import ctypes as C
    det = capi.detector("LUX_Run03")
    rc = capi.RunConsts()
    assert orc.L.orc_prepare(C.byref(det), C.c_double(180.0), C.byref(rc)) == 0
    c = gold["consts_180"]
    got = [rc.rho, rc.Wq_eV, rc.alpha, rc.vD] + list(rc.g2_params) + [rc.NewMean]
    assert got == list(c)
    x, y, z = gold["maps_xyz"]
    assert np.array_equal(np.stack(orc.maps(x, y, z)), gold["maps"])


def test_chain_bitwise(orc, gold):
    cases = ast.literal_eval(str(gold["chain_cases"]))
    for name, (seed, n, sp, kind, e0, e1, fld, fp, xyz, erm) in cases.items():
        got = ref_chain(orc, seed, n, sp, kind, e0, e1, fld, fp, xyz, erm)
        assert np.array_equal(got, gold[name]), name


def test_band_samplers_spectra_lar_vec_bitwise(orc, gold):
    ana = capi.analysis(0)
    s1, s2 = gold["band_sig"]
    assert np.array_equal(orc.getband(s1, s2, ana=ana), gold["band"], equal_nan=True)
    for key in gold.files:
        if key.startswith("sampler_"):
            _, kind, p0, p1 = key.split("_")
            p2 = {1: 1.0, 3: 2.25, 5: 1000.0}.get(int(kind), 0.0)
            assert np.array_equal(orc.sampler(int(kind), 5, 400, float(p0), float(p1), p2), gold[key]), key
    for k, nme in [(capi.SRC_CH3T, "CH3T"), (capi.SRC_DD, "DD"), (capi.SRC_B8, "B8")]:
        assert np.array_equal(orc.spectrum(k, 9, 400, 0.0, 80.0), gold["spectrum_" + nme])
    for k, nme, hi in [(capi.SRC_C14, "C14", 156.0), (capi.SRC_CF, "Cf", 200.0), (capi.SRC_PPSOLAR, "ppSolar", 250.0),
                       (capi.SRC_ATMNU, "atmNu", 85.0), (capi.SRC_AMBE, "AmBe", 300.0)]:
        assert np.array_equal(orc.spectrum(k, 9, 400, 0.5, hi), gold["spectrum_" + nme]), nme
    for sp in (0, 1, 2):
        for fld in (1.0, 500.0, 1000.0):
            y, f = orc.lar(4, sp, gold["lar_E"], fld, 1.393)
            assert np.array_equal(y, gold["lar_y_%d_%g" % (sp, fld)])
            assert np.array_equal(f, gold["lar_f_%d_%g" % (sp, fld)])
    m0 = capi.model(0)
    args = (list(m0.ERYieldsParam), list(m0.NRYieldsParam), list(m0.NRERWidthsParam)[:13])
    assert np.array_equal(orc.runvec(capi.beta, gold["vec_E"], gold["vec_xyz"], 180.0, 3, *args), gold["vec_beta_180"])
    assert np.array_equal(orc.runvec(capi.NR, gold["vec_E"], gold["vec_xyz"], -1.0, 3, *args), gold["vec_NR_fitEF"])


def test_survey_pinned_values(orc):
    """values captured from the compiled reference in SURVEY.md 8(c)"""
    mo = capi.model(1)
    nr, er = list(mo.NRYieldsParam), list(mo.ERYieldsParam)
    b = orc.yields(1, capi.beta, [5.0, 0.5, 20.0, 100.0], [180.0, 180.0, 50.0, 1000.0], RHO_LUX, 131.293, 54., nr, er)
    assert list(b[0][:3]) == [205.88604645954106, 165.20770456738973, 0.050298182684348262]
    assert list(b[3][:3]) == [2806.3962674277627, 4615.4787531108532, 0.18202453502051572]
    n = orc.yields(0, capi.NR, [1.0, 10.0], [180.0, 180.0], RHO_LUX, 131.293, 54., nr, er)
    assert list(n[0][:4]) == [3.8175326853690859, 6.8511840288214803, 0.50923264205705066, 0.14374691954077559]
    g = orc.yields(0, capi.gammaRay, [41.5], [180.0], RHO_LUX, 131.293, 54., nr, er)
    assert list(g[0][:3]) == [2255.6704922630879, 824.40764126043757, 0.18141640390708169]
    y, _ = orc.lar(1, 0, [10.0], 500.0, 1.393)
    assert (y[0][3], y[0][4]) == (86.11970082647197, 49.498412045124624)


# ---- live reference (this container; oracle/_ref travels to the GPU box as prebuilt files) ------

def test_oracle_equals_live_reference(orc, ref):
    rng = np.random.default_rng(0)
    for seed in (17, 18):
        for sp, kind, e0, e1, erm in [(capi.beta, capi.SRC_FLAT, 0.1, 20.0, 0), (capi.NR, capi.SRC_FLAT, 0.0, 100.0, -1),
                                      (capi.beta, capi.SRC_GAUSS, 10.0, 2.0, 1), (capi.beta, capi.SRC_FLAT, 0.5, 60.0, 2),
                                      (capi.Kr83m, capi.SRC_MONO, 41.5, 600.0, -1), (capi.C14, capi.SRC_C14, 0., 156., -1),
                                      (capi.atmNu, capi.SRC_ATMNU, 0.0, 85.0, -1)]:
            a = ref_chain(ref, seed, 1500, sp, kind, e0, e1, 180.0, 0, (0, 0, 0), erm, 1.2)
            b = ref_chain(orc, seed, 1500, sp, kind, e0, e1, 180.0, 0, (0, 0, 0), erm, 1.2)
            assert np.array_equal(a, b), (seed, sp, kind)
    E = rng.uniform(0.01, 300.0, 3000)
    F = rng.uniform(1.0, 3000.0, 3000)
    mo = capi.model(0)
    for which, sp in [(0, capi.NR), (1, capi.beta), (2, capi.beta), (3, capi.beta)]:
        assert np.array_equal(ref.yields(which, sp, E, F, 2.95, 131.293, 54., list(mo.NRYieldsParam),
                                         list(mo.ERYieldsParam), 0.8),
                              orc.yields(which, sp, E, F, 2.95, 131.293, 54., list(mo.NRYieldsParam),
                                         list(mo.ERYieldsParam), 0.8))


def test_oracle_reproduces_reference_binary_stdout(orc):
    """`execNEST 300 ER 0.1 20 180 -1 1` of the unmodified LUX_Run03 reference binary, row by row
    (verbosity 1: CalculateG2's 10 000-sample Monte Carlo consumes the stream before event 0)"""
    if not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("oracle/_ref/execNEST_ref_lux not built")
    n = 300
    r = subprocess.run([refprobe.REF_BIN_LUX, str(n), "ER", "0.1", "20", "180", "-1", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.DEVNULL, text=True, check=True)
    rows = [ln for ln in r.stdout.splitlines() if ln[:1].isdigit() or ln[:1] == "-"]
    rows = [ln for ln in rows if ln.count("\t") >= 11]
    assert len(rows) == n
    mo = capi.model(1)
    o = orc.chain(1, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, 180.0, 0, [0., 0., 0.], 0, 1.0, list(mo.NRYieldsParam),
                  list(mo.ERYieldsParam), list(mo.NRERWidthsParam)[:13], 0.0, 1, ana=capi.analysis(0))
    C = refprobe.COL
    ana = capi.analysis(0)
    for j, ln in enumerate(rows):
        s1, s2 = o[j, refprobe.S1_0:refprobe.S1_0 + 9], o[j, refprobe.S2_0:refprobe.S2_0 + 9]
        exp = "%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d\t" % (
            o[j, C["keV_out"]], o[j, C["field"]], o[j, C["dt"]], o[j, C["sx"]], o[j, C["sy"]], o[j, C["z"]],
            o[j, C["Nph"]], o[j, C["Ne"]])
        # execNEST.cpp:1390-1408: engineering notation once a pulse leaves the analysis window
        if o[j, C["keV_out"]] > 1000. or s1[5] > ana.maxS1 or s2[7] > ana.maxS2:
            exp += "%e\t%e\t%e\t%d\t%e\t%e" % (s1[2], s1[5], s1[7], int(s2[0]), s2[4], s2[7])
        else:
            exp += "%.6f\t%.6f\t%.6f\t%i\t%.6f\t%.6f" % (s1[2], s1[5], s1[7], int(s2[0]), s2[4], s2[7])
        assert ln.strip() == exp, (j, ln, exp)


def test_oracle_stage_calls_equal_live_reference(orc, ref):
    """GetQuanta, GetS1 and GetS2 on fixed inputs -- the inputs of the GPU stage-call tests -- bit for bit (same seed,
    same generator, same draw order)"""
    import test_gpu_zz_stage_api as cases
    mo = capi.model(1)
    widths = list(mo.NRERWidthsParam)[:mo.n_widths]
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    for name, (y6, skew) in cases.RECORDS.items():
        ai, ad = ref.quanta(5, 3000, y6, RHO_LUX, widths, skew)
        bi, bd = orc.quanta(5, 3000, y6, RHO_LUX, widths, skew)
        assert np.array_equal(ai, bi) and np.array_equal(ad, bd), name
    for name, (Nph, pos, spos, energy, mode) in cases.S1_CASES.items():
        n = 3000 if Nph < 5000 else 500
        a = ref.s1(6, n, Nph, pos, spos, rc.vD, rc.vD_middle, capi.beta, 180.0, energy, mode)
        b = orc.s1(6, n, Nph, pos, spos, rc.vD, rc.vD_middle, capi.beta, 180.0, energy, mode)
        assert np.array_equal(a, b), name
    for name, (Ne, pos, spos, dt, field) in cases.S2_CASES.items():
        a = ref.s2(7, 3000, Ne, pos, spos, dt, rc.vD, field)
        b = orc.s2(7, 3000, Ne, pos, spos, dt, rc.vD, field)
        assert np.array_equal(a, b), name


def test_env_configurable_reference_binary_is_the_reference_when_nothing_is_set():
    """oracle/_ref/execNEST_ref_lux_env (analysis globals settable from NEST_REF_* variables, oracle/shim_lux_env) must be
    execNEST_ref_lux when no variable is set -- byte for byte -- and must obey a variable when one is"""
    import os
    import subprocess
    import refprobe
    if not (os.path.exists(refprobe.REF_BIN_LUX) and os.path.exists(refprobe.REF_BIN_LUX_ENV)):
        pytest.skip("reference binaries not built")
    args = ["300", "ER", "0.1", "20", "180", "-1", "1"]
    env = {k: v for k, v in os.environ.items() if not k.startswith("NEST_REF_")}
    a = subprocess.run([refprobe.REF_BIN_LUX] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    b = subprocess.run([refprobe.REF_BIN_LUX_ENV] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert a.returncode == b.returncode == 0
    assert a.stdout == b.stdout and a.stderr == b.stderr
    env["NEST_REF_USEPD"] = "0"
    c = subprocess.run([refprobe.REF_BIN_LUX_ENV] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert c.returncode == 0 and c.stdout != a.stdout                      # areas in phe: the energy column moves
    rows_a = [ln.split(b"\t") for ln in a.stdout.splitlines() if ln[:1].isdigit()]
    rows_c = [ln.split(b"\t") for ln in c.stdout.splitlines() if ln[:1].isdigit()]
    assert len(rows_a) == len(rows_c) == 300
    assert all(x[1:] == y[1:] for x, y in zip(rows_a, rows_c))             # same events: everything but E_recon equal


@pytest.mark.parametrize("usePD,useS2,tE,tP,win", [(0, 0, 0, 0, (2.0, 165.0, 80.0, 2.0e4)), (1, 0, 0, 0, (1.5, 120.0, 60.0, 1.5e4)),
                                                    (2, 1, 1, 0, (1.5, 99.5, 42.0, 1.0e4)), (1, 2, 0, 1, (1.5, 120.0, 60.0, 1.5e4)),
                                                    (0, 1, 0, 1, (2.0, 165.0, 80.0, 2.0e4))])
@pytest.mark.parametrize("species,tname,emin,emax", [(8, "beta", 0.1, 20.0), (0, "NR", 0.5, 60.0)])
def test_oracle_reproduces_reference_binary_in_other_analysis_modes(orc, usePD, useS2, tE, tP, win, species, tname, emin, emax):
    """the restatement's signal selection, energy reconstruction (execNEST.cpp:1121-1250), MC-truth switches and GetBand in
    the analysis modes the shipped settings do not reach, against the reference binary with the same switches
    (execNEST_ref_lux_env): the printed energy and position of every row and the whole band table, digit for digit"""
    if not os.path.exists(refprobe.REF_BIN_LUX_ENV):
        pytest.skip("oracle/_ref/execNEST_ref_lux_env not built")
    n = 4000
    env = {k: v for k, v in os.environ.items() if not k.startswith("NEST_REF_")}
    env.update({"NEST_REF_USEPD": str(usePD), "NEST_REF_USES2": str(useS2), "NEST_REF_MCTRUTHE": str(tE),
                "NEST_REF_MCTRUTHPOS": str(tP), "NEST_REF_MINS1": repr(win[0]), "NEST_REF_MAXS1": repr(win[1]),
                "NEST_REF_MINS2": repr(win[2]), "NEST_REF_MAXS2": repr(win[3])})
    if useS2 == 1:
        env.update({"NEST_REF_LOGMIN": "1.5", "NEST_REF_LOGMAX": "4.5"})
    r = subprocess.run([refprobe.REF_BIN_LUX_ENV, str(n), tname, repr(emin), repr(emax), "180", "-1", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True, env=env)
    rows = [ln for ln in r.stdout.splitlines() if (ln[:1].isdigit() or ln[:1] == "-") and ln.count("\t") >= 11]
    assert len(rows) == n
    ana = capi.analysis(0)
    ana.usePD, ana.useS2, ana.MCtruthE, ana.MCtruthPos = usePD, useS2, tE, tP
    ana.minS1, ana.maxS1, ana.minS2, ana.maxS2 = win
    if useS2 == 1:
        ana.logMin, ana.logMax = 1.5, 4.5
    mo = capi.model(1)
    o = orc.chain(1, n, species, capi.SRC_FLAT, emin, emax, 180.0, 0, [0., 0., 0.], -1, 1.0, list(mo.NRYieldsParam),
                  list(mo.ERYieldsParam), list(mo.NRERWidthsParam)[:13], 0.0, 1, ana=ana)
    C = refprobe.COL
    for j, ln in enumerate(rows):
        f = ln.split("\t")
        # energy as printed (reconstructed in this mode's units, or the true one), position (smeared, or the true one)
        xs, ys = (o[j, C["x"]], o[j, C["y"]]) if tP else (o[j, C["sx"]], o[j, C["sy"]])
        exp = "%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d" % (o[j, C["keV_out"]], o[j, C["field"]], o[j, C["dt"]], xs, ys,
                                                          o[j, C["z"]], o[j, C["Nph"]], o[j, C["Ne"]])
        assert "\t".join(f[:6]) == exp, (j, ln, exp)
    # the band table on stderr against GetBand (restated) on the restatement's own signals
    band = []
    for ln in r.stderr.splitlines():
        g = ln.split()
        if len(g) == 7:
            try:
                band.append([float(v) for v in g])
            except ValueError:
                pass
    band = np.array(band)
    a1, a2 = (refprobe.SIG2, refprobe.SIG1) if useS2 == 2 else (refprobe.SIG1, refprobe.SIG2)
    ob = orc.getband(o[:, a1], o[:, a2], ana=ana)
    assert band.shape == (ana.numBins, 7) and len(ob) == ana.numBins
    # printed columns: centre, actual, mean, mean error, sigma, skew, eff[%] <- GetBand's 0, 1, 2, 4, 3, 6, 5 * 100
    exp = np.column_stack([ob[:, 0], ob[:, 1], ob[:, 2], ob[:, 4], ob[:, 3], ob[:, 6], ob[:, 5] * 100.0])
    both = np.isfinite(band) & np.isfinite(exp)
    assert np.array_equal(np.isnan(band), np.isnan(exp))
    assert np.allclose(band[both], exp[both], rtol=0, atol=6e-7)           # "%lf": six decimals


VARIANTS = {   # oracle/shim_variants/LUX_variants.hh, as changes to the flattened LUX_Run03
    "luxbot3": dict(s2_thr=-150.0, coinLevel=3),
    "luxnoisy": dict(sPEeff=0.85, sPEthr=0.5, P_dphe=0.25, noiseLinear=(3.e-2, 2.e-2), noiseQuadratic=(1.e-4, 2.e-5),
                     noiseBaseline=(0.02, 0.1, 0.0, 0.0)),
    "luxnocoin": dict(coinLevel=0, eLife_us=250.0, E_gas=5.0, s2Fano=1.5, coinWind=50.0),
    "luxwarm": dict(T_Kelvin=170.0, p_bar=1.7, OldW13eV=0, PosResFlat=5.0, PosResBase=120.0, coinLevel=1),
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("species,tname,emin,emax", [(8, "ER", 0.1, 20.0), (0, "NR", 0.5, 60.0)])
def test_oracle_reproduces_reference_binary_for_user_detectors(orc, variant, species, tname, emin, emax):
    """the restatement's detector-parameter branches -- bottom-array S2, 3-fold / no coincidence, sPEeff < 1, linear,
    quadratic and baseline noise, other lifetime / extraction field / S2 Fano factor / (T, p, W) -- against the reference
    binary built with the same LUX-derived user detector (oracle/_ref/execNEST_ref_<variant>): rows digit for digit"""
    path = os.path.join(os.path.dirname(refprobe.REF_BIN_LUX), "execNEST_ref_" + variant)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/execNEST_ref_%s not built" % variant)
    n = 300
    r = subprocess.run([path, str(n), tname, repr(emin), repr(emax), "180", "-1", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.DEVNULL, text=True, check=True)
    rows = [ln for ln in r.stdout.splitlines() if (ln[:1].isdigit() or ln[:1] == "-") and ln.count("\t") >= 11]
    assert len(rows) == n
    saved = capi.Detector.from_buffer_copy(orc.det)
    try:
        for k, v in VARIANTS[variant].items():
            if isinstance(v, tuple):
                for i, x in enumerate(v):
                    getattr(orc.det, k)[i] = x
            else:
                setattr(orc.det, k, v)
        capi.prepare(orc.det, 180.0)
        mo = capi.model(1)
        er_model = 0 if tname == "ER" else -1
        ana = capi.analysis(0)
        o = orc.chain(1, n, species, capi.SRC_FLAT, emin, emax, 180.0, 0, [0., 0., 0.], er_model, 1.0, list(mo.NRYieldsParam),
                      list(mo.ERYieldsParam), list(mo.NRERWidthsParam)[:13], 0.0, 1, ana=ana)
    finally:
        C.memmove(C.byref(orc.det), C.byref(saved), C.sizeof(saved))
    Cc = refprobe.COL
    for j, ln in enumerate(rows):
        s1, s2 = o[j, refprobe.S1_0:refprobe.S1_0 + 9], o[j, refprobe.S2_0:refprobe.S2_0 + 9]
        exp = "%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d\t" % (
            o[j, Cc["keV_out"]], o[j, Cc["field"]], o[j, Cc["dt"]], o[j, Cc["sx"]], o[j, Cc["sy"]], o[j, Cc["z"]],
            o[j, Cc["Nph"]], o[j, Cc["Ne"]])
        if o[j, Cc["keV_out"]] > 1000. or s1[5] > ana.maxS1 or s2[7] > ana.maxS2:
            exp += "%e\t%e\t%e\t%d\t%e\t%e" % (s1[2], s1[5], s1[7], int(s2[0]), s2[4], s2[7])
        else:
            exp += "%.6f\t%.6f\t%.6f\t%i\t%.6f\t%.6f" % (s1[2], s1[5], s1[7], int(s2[0]), s2[4], s2[7])
        assert ln.strip() == exp, (variant, j, ln, exp)
