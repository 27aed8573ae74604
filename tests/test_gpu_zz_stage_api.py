"""The per-event stage call GetQuanta (reference src/NEST.cpp:248-432; FullCalculation :17-40, bareNEST.cpp:186)
through the C-ABI (nb200_quanta -> k_quanta) against the unmodified reference: n draws on fixed YieldResults,
every QuantaResult field compared -- recombProb to 1e-12 (it is deterministic given the yields), the integer
quanta and Variance as distributions -- plus conservation, independence of how a list is split over calls,
and the reference's error for fewer than 13 width parameters."""
import numpy as np
import pytest
from scipy import stats

from helpers import RHO_LUX
from nest_b200 import capi

pytestmark = pytest.mark.gpu

# YieldResults of the reference at 180 V/cm: {Nph, Ne, NexONi, Lindhard, field, DeltaT_Scint}
RECORDS = {
    "beta 5 keV": ([205.88604645954106, 165.20770456738973, 0.050298182684348262, 1.0, 180.0, -999.0], 0.0),
    "beta 5 keV, LUX skew model": ([205.88604645954106, 165.20770456738973, 0.050298182684348262, 1.0, 180.0, -999.0], -999.0),
    "beta 0.5 keV": ([0.64889386770322943, 36.460481234989849, 0.0051337477773812801, 1.0, 180.0, -999.0], 0.0),
    "gamma 41.5 keV": ([2255.6704922630879, 824.40764126043757, 0.18141640390708169, 1.0, 180.0, -999.0], 0.0),
    "NR 10 keV": ([80.68553759012174, 57.878421270493256, 0.81936822679002697, 0.18669670194818133, 180.0, -999.0], 0.0),
    "NR 0.3 keV": ([0.88939281415746496, 1.1501592836462304, 0.76402647275054425, 0.091600936068923555, 180.0, -999.0], 0.0),
    "alpha 5.3 MeV": ([368873.82323330763, 5269.1361310607754, 0.0053475469399098521, 0.91499960874930175, 180.0, -999.0], 0.0),
}


@pytest.mark.parametrize("name", list(RECORDS))
def test_quanta_stage_matches_reference(ctx, ref, name):
    y6, skew = RECORDS[name]
    n = 20000
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)                       # the CLI's width parameters (execNEST.cpp:188-204)
    mo.SkewnessER = skew
    widths = list(mo.NRERWidthsParam)[:mo.n_widths]
    qi, qd = ctx.quanta(det, mo, np.tile(np.array(y6)[:, None], (1, n)), RHO_LUX, seed=11)
    ri, rd = ref.quanta(12, n, y6, RHO_LUX, widths, skew)
    # conservation (the reference throws otherwise, NEST.cpp:421-424) and the clamps
    assert np.all(qi[0] + qi[1] == qi[2] + qi[3])
    assert np.all(qi >= 0) and np.all(qi[1] <= qi[2]) and np.all(qi[0] >= qi[3])
    # recombProb is a function of the yields alone wherever quanta were made
    made = qd[0] != 0
    rmade = rd[:, 0] != 0
    assert abs(made.mean() - rmade.mean()) < 5 * np.sqrt(0.25 / n) + 1e-12
    if made.any():
        assert np.unique(qd[0][made]).size == 1 and np.unique(rd[rmade, 0]).size == 1
        assert qd[0][made][0] == pytest.approx(rd[rmade, 0][0], rel=1e-12)
    for k, col in enumerate(("photons", "electrons", "ions", "excitons")):
        a, b = qi[k].astype(float), ri[:, k].astype(float)
        if np.all(a == a[0]) and np.all(b == b[0]):
            assert a[0] == b[0], col
            continue
        assert stats.ks_2samp(a, b).pvalue > 1e-4, (name, col, a.mean(), b.mean())
        se = np.sqrt(a.var() / n + b.var() / n)
        assert abs(a.mean() - b.mean()) < 5 * se + 1e-9, (name, col, a.mean(), b.mean())
    a, b = qd[1], rd[:, 1]
    if not (np.all(a == a[0]) and np.all(b == b[0])):
        assert stats.ks_2samp(a, b).pvalue > 1e-4, (name, "Variance", a.mean(), b.mean())
    # Variance = lambda r (1 - r) Ni + omega^2 Ni^2 (:354-355) is a function of Ni alone for fixed yields: wherever the two
    # sides drew the same Ni they must report the same Variance
    lut = {int(a): b for a, b in zip(ri[rmade, 2], rd[rmade, 1])}
    pairs = [(qd[1][i], lut[int(qi[2][i])]) for i in np.nonzero(made)[0] if int(qi[2][i]) in lut]
    assert len(pairs) > 0.5 * made.sum()
    assert max(abs(a - b) / b for a, b in pairs) < 1e-10


def test_quanta_stage_does_not_depend_on_batching(ctx):
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.massNum, mo.er_model = 131.293, 0
    E = np.linspace(0.2, 60.0, 5000)
    y = ctx.yields(det, mo, capi.beta, E, 180.0, RHO_LUX)
    mo.er_model = -1
    whole_i, whole_d = ctx.quanta(det, mo, y, RHO_LUX, seed=3)
    a_i, a_d = ctx.quanta(det, mo, y[:, :1234], RHO_LUX, seed=3)
    b_i, b_d = ctx.quanta(det, mo, y[:, 1234:], RHO_LUX, seed=3, first_event=1234)
    assert np.array_equal(np.concatenate([a_i, b_i], axis=1), whole_i)
    assert np.array_equal(np.concatenate([a_d, b_d], axis=1), whole_d)
    other_i, _ = ctx.quanta(det, mo, y, RHO_LUX, seed=4)
    assert not np.array_equal(other_i, whole_i)
    # mean quanta follow the mean yields
    assert abs((whole_i[0] + whole_i[1]).mean() / (y[0] + y[1]).mean() - 1) < 0.01


def test_quanta_stage_errors(ctx):
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.n_widths = 12                          # NEST.cpp:258-262: "You need a minimum of 13 free parameters ..."
    y = np.tile(np.array(RECORDS["beta 5 keV"][0])[:, None], (1, 4))
    with pytest.raises(capi.NestB200Error) as e:
        ctx.quanta(det, mo, y, RHO_LUX, seed=1)
    assert e.value.code == capi.EPHYSICS and "13 free parameters" in str(e.value)
    with pytest.raises(ValueError):
        ctx.quanta(det, capi.model(1), y[:5], RHO_LUX, seed=1)


# GetS2 on fixed inputs: (electrons, truth xyz, smeared xyz, drift time [us], field [V/cm])
S2_CASES = {
    "165 e-": (165, (50., -30., 250.), (52., -28., 250.), 150.0, 180.0),
    "5 e-": (5, (0., 0., 300.), (0., 0., 300.), 100.0, 180.0),
    "4000 e- near the wall, long drift": (4000, (120., 80., 100.), (118., 83., 100.), 300.0, 180.0),
    "no electrons": (0, (10., 10., 200.), (10., 10., 200.), 120.0, 180.0),
    "no field": (300, (10., 10., 200.), (10., 10., 200.), 120.0, 0.5),
}


@pytest.mark.parametrize("name", list(S2_CASES))
def test_s2_stage_matches_reference(ctx, ref, name):
    """NESTcalc::GetS2 Full (NEST.cpp:1724-2063; bareNEST.cpp:210) through nb200_s2 -> k_s2, all nine ionization[]
    values against the unmodified reference on the same fixed inputs"""
    Ne, pos, spos, dt, field = S2_CASES[name]
    n = 20000
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    p5 = np.tile(np.array([pos[0], pos[1], pos[2], spos[0], spos[1]])[:, None], (1, n))
    got = ctx.s2(det, rc, np.full(n, Ne), p5, dt, field, seed=21)
    exp = ref.s2(22, n, Ne, pos, spos, dt, rc.vD, field).T
    names = ["Nee", "Nph", "nHits", "Nphe", "area", "areaC", "phd", "phdC", "g2"]
    for k, col in enumerate(names):
        a, b = got[k], exp[k]
        if np.all(a == a[0]) and np.all(b == b[0]):
            assert a[0] == pytest.approx(b[0], rel=1e-12, abs=0), (name, col, a[0], b[0])
            continue
        assert stats.ks_2samp(a, b).pvalue > 1e-4, (name, col, a.mean(), b.mean())
        se = np.sqrt(a.var() / n + b.var() / n)
        assert abs(a.mean() - b.mean()) < 5 * se + 1e-9, (name, col, a.mean(), b.mean())
    # the sign flag (below |s2_thr|: the first eight negated, NEST.cpp:2051-2057) at the same rate
    assert abs((got[4] < 0).mean() - (exp[4] < 0).mean()) < 5 * np.sqrt(0.25 / n) + 1e-12


def test_s2_stage_does_not_depend_on_batching(ctx):
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    rng = np.random.default_rng(1)
    n = 6000
    Ne = rng.integers(0, 3000, n)
    r, phi = 200 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(50, 500, n)
    pos = np.array([r * np.cos(phi), r * np.sin(phi), z, r * np.cos(phi) + 1.0, r * np.sin(phi) - 1.0])
    dt = (det.TopDrift - z) / rc.vD
    whole = ctx.s2(det, rc, Ne, pos, dt, 180.0, seed=5)
    a = ctx.s2(det, rc, Ne[:2500], np.ascontiguousarray(pos[:, :2500]), dt[:2500], 180.0, seed=5)
    b = ctx.s2(det, rc, Ne[2500:], np.ascontiguousarray(pos[:, 2500:]), dt[2500:], 180.0, seed=5, first_event=2500)
    assert np.array_equal(np.concatenate([a, b], axis=1), whole)
    assert not np.array_equal(ctx.s2(det, rc, Ne, pos, dt, 180.0, seed=6), whole)
    # extracted electrons follow the lifetime and the extraction efficiency (NEST.cpp:1775)
    expect = (Ne * rc.g2_params[1] * np.exp(-dt / det.eLife_us)).sum()
    assert abs(np.abs(whole[0]).sum() / expect - 1) < 0.01
    with pytest.raises(ValueError):
        ctx.s2(det, rc, Ne, pos[:4], dt, 180.0, seed=5)


# GetS1 on fixed inputs: (photons, truth xyz, smeared xyz, energy [keV], S1 mode)
S1_CASES = {
    "2000 photons, Full": (2000, (50., -30., 250.), (52., -28., 250.), 10.0, 0),
    "20 photons, Full (coincidence level matters)": (20, (0., 0., 300.), (1., 1., 300.), 0.5, 0),
    "no photons": (0, (10., 10., 200.), (10., 10., 200.), 1.0, 0),
    "30000 photons, Hybrid at 150 keV (Parametric)": (30000, (120., 80., 100.), (118., 83., 100.), 150.0, 2),
    "300 photons, Parametric": (300, (100., -100., 450.), (98., -97., 450.), 5.0, 1),
}


@pytest.mark.parametrize("name", list(S1_CASES))
def test_s1_stage_matches_reference(ctx, ref, name):
    """NESTcalc::GetS1 (NEST.cpp:1288-1722; bareNEST.cpp:203) through nb200_s1 -> k_s1pre, k_hits, k_s1post: all nine
    scintillation[] values against the unmodified reference on the same fixed inputs"""
    Nph, pos, spos, energy, mode = S1_CASES[name]
    n = 20000 if Nph < 5000 else 5000
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    p6 = np.tile(np.array(list(pos) + list(spos))[:, None], (1, n))
    got = ctx.s1(det, rc, np.full(n, Nph), p6, energy, seed=31, mode=mode)
    exp = ref.s1(32, n, Nph, pos, spos, rc.vD, rc.vD_middle, capi.beta, 180.0, energy, mode).T
    names = ["nHits", "Nphe", "area", "areaC", "phd", "phdC", "spike", "spikeC", "spike count"]
    for k, col in enumerate(names):
        a, b = got[k], exp[k]
        if np.all(a == a[0]) and np.all(b == b[0]):
            assert a[0] == pytest.approx(b[0], rel=1e-12, abs=0), (name, col, a[0], b[0])
            continue
        assert stats.ks_2samp(a, b).pvalue > 1e-4, (name, col, a.mean(), b.mean())
        se = np.sqrt(a.var() / n + b.var() / n)
        assert abs(a.mean() - b.mean()) < 5 * se + 1e-9, (name, col, a.mean(), b.mean())
    # the coincidence flag (all eight negated, NEST.cpp:1689-1716) at the same rate
    assert abs((got[2] < 0).mean() - (exp[2] < 0).mean()) < 5 * np.sqrt(0.25 / n) + 1e-12


def test_s1_stage_does_not_depend_on_batching(ctx):
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    rng = np.random.default_rng(2)
    n = 6000
    Nph = rng.integers(0, 4000, n)
    r, phi = 200 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(50, 500, n)
    x, y = r * np.cos(phi), r * np.sin(phi)
    pos = np.array([x, y, z, x + 1.0, y - 1.0, z])
    E = rng.uniform(0.5, 200.0, n)                       # Hybrid: Full below 100 keV, Parametric above
    whole = ctx.s1(det, rc, Nph, pos, E, seed=5)
    a = ctx.s1(det, rc, Nph[:2500], np.ascontiguousarray(pos[:, :2500]), E[:2500], seed=5)
    b = ctx.s1(det, rc, Nph[2500:], np.ascontiguousarray(pos[:, 2500:]), E[2500:], seed=5, first_event=2500)
    assert np.array_equal(np.concatenate([a, b], axis=1), whole)
    assert not np.array_equal(ctx.s1(det, rc, Nph, pos, E, seed=6), whole)
    # hits ~ g1 x light-collection map x photons: within the map's range of the detector's g1
    ratio = np.abs(whole[0]).sum() / Nph.sum()
    assert 0.6 * det.g1 < ratio < 1.4 * det.g1
    with pytest.raises(capi.NestB200Error):
        ctx.s1(det, rc, Nph, pos, E, seed=5, mode=3)     # Waveform: not part of this path


@pytest.mark.parametrize("wset", ["default", "perturbed_single", "perturbed_double"])
@pytest.mark.parametrize("name,skew", [("beta 5 keV", -999.0), ("beta 5 keV", 1.5), ("gamma 41.5 keV", -999.0),
                                        ("NR 10 keV", 0.0), ("alpha 5.3 MeV", 0.0)])
def test_quanta_stage_with_other_width_parameters(ctx, ref, wset, name, skew):
    """GetQuanta with the width parameters other than the command line's: the library defaults (NEST.hh:122-123, [8] < 1:
    RecombOmegaER's single skew-Gaussian in field and energy, NEST.cpp:173-196) and user-fitted sets in both regimes of
    [8], with SkewnessER = -999 (the LUX skew model of :379-388) and a fixed skewness -- against the reference"""
    if name not in RECORDS:
        pytest.skip("record not in this file's table")
    y6, _ = RECORDS[name]
    n = 20000
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(0 if wset == "default" else 1)
    w = np.array(list(mo.NRERWidthsParam)[:mo.n_widths])
    rng = np.random.default_rng(7)
    if wset == "perturbed_single":
        w = np.array(list(capi.model(0).NRERWidthsParam)[:13]) * (1.0 + 0.2 * rng.uniform(-1, 1, 13))
        w[8] = 0.6                                   # < 1: single-Gaussian branch; [9] must stay below it (:255-262 guards)
        w[9] = 0.25
    elif wset == "perturbed_double":
        w = w * (1.0 + 0.1 * rng.uniform(-1, 1, w.size))
        w[8] = 2.2                                   # >= 1: the double Gaussian in log10(Nq) (:197-222)
    mo.n_widths = len(w)
    for i, v in enumerate(w):
        mo.NRERWidthsParam[i] = v
    mo.SkewnessER = skew
    qi, qd = ctx.quanta(det, mo, np.tile(np.array(y6)[:, None], (1, n)), RHO_LUX, seed=21)
    ri, rd = ref.quanta(22, n, y6, RHO_LUX, list(w), skew)
    assert np.all(qi[0] + qi[1] == qi[2] + qi[3])
    made, rmade = qd[0] != 0, rd[:, 0] != 0
    if made.any() and rmade.any():
        assert qd[0][made][0] == pytest.approx(rd[rmade, 0][0], rel=1e-12)          # recombination probability
    for k, col in enumerate(("photons", "electrons", "ions", "excitons")):
        a, b = qi[k].astype(float), ri[:, k].astype(float)
        if np.all(a == a[0]) and np.all(b == b[0]):
            assert a[0] == b[0], col
            continue
        assert stats.ks_2samp(a, b).pvalue > 1e-4, (wset, name, skew, col, a.mean(), b.mean(), a.std(), b.std())
    lut = {int(a): b for a, b in zip(ri[rmade, 2], rd[rmade, 1])}
    pairs = [(qd[1][i], lut[int(qi[2][i])]) for i in np.nonzero(made)[0] if int(qi[2][i]) in lut]
    if pairs:
        assert max(abs(a - b) / max(abs(b), 1e-300) for a, b in pairs) < 1e-10                  # Variance(Ni)
