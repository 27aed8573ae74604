"""Deterministic mean yields on the GPU against the compiled reference: 1e-12 relative
(BASELINE.json north_star).  Reference: NESTcalc::GetYields src/NEST.cpp:1211-1271."""
import numpy as np
import pytest

from helpers import RHO_LUX
from nest_b200 import capi

pytestmark = pytest.mark.gpu
TOL = 1e-12


def grid(emin, emax, ne, fields):
    E = np.exp(np.linspace(np.log(emin), np.log(emax), ne))
    EE, FF = np.meshgrid(E, np.asarray(fields, float), indexing="ij")
    return EE.ravel().copy(), FF.ravel().copy()


def relerr(a, b):
    denom = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(denom > 0, np.abs(a - b) / denom, 0.0)
    return r


FIELDS = [1.0, 10.0, 50.0, 180.0, 400.0, 1000.0, 4000.0]


@pytest.mark.parametrize("name,species,er_model,which_ref,mult,erange", [
    ("NR", capi.NR, -1, 0, 1.0, (0.05, 330.0)),
    ("DD", capi.DD, -1, 0, 1.0, (0.05, 80.0)),
    ("beta", capi.beta, -1, 0, 1.0, (0.01, 3000.0)),
    ("betaGR_mult", capi.beta, 0, 1, 0.9, (0.01, 3000.0)),
    ("gamma", capi.gammaRay, -1, 0, 1.0, (0.01, 3000.0)),
    ("gamma_mult", capi.beta, 1, 2, 1.3, (0.01, 3000.0)),
    ("ERweighted", capi.beta, 2, 3, 1.0, (0.01, 3000.0)),
])
def test_yields_match_reference(ctx, ref, name, species, er_model, which_ref, mult, erange):
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.er_model, mo.multFact, mo.massNum, mo.atomNum = er_model, mult, 131.293, 54.0
    E, F = grid(erange[0], erange[1], 400, FIELDS)
    for rho in (RHO_LUX, 2.7, 3.05):
        got = ctx.yields(det, mo, species, E, F, rho)
        exp = ref.yields(which_ref, species, E, F, rho, 131.293, 54.0, list(mo.NRYieldsParam),
                         list(mo.ERYieldsParam), mult).T
        for k, col in enumerate(["Nph", "Ne", "NexONi", "L", "field", "dT"]):
            r = relerr(got[k], exp[k])
            assert r.max() <= TOL, "%s %s rho=%g: max rel err %.3e at E=%g F=%g (gpu %r ref %r)" % (
                name, col, rho, r.max(), E[r.argmax()], F[r.argmax()], got[k][r.argmax()], exp[k][r.argmax()])
        # a flipped validity/clamp branch would be an O(1) difference: zeros must coincide exactly
        assert np.array_equal(got[0] == 0, exp[0] == 0) and np.array_equal(got[1] == 0, exp[1] == 0)


def test_alpha_yields(ctx, ref):
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.massNum, mo.atomNum = 4.0, 2.0
    E, F = grid(100.0, 9000.0, 200, FIELDS)
    got = ctx.yields(det, mo, capi.ion, E, F, RHO_LUX)
    exp = ref.yields(0, capi.ion, E, F, RHO_LUX, 4.0, 2.0, list(mo.NRYieldsParam), list(mo.ERYieldsParam)).T
    for k in range(4):
        assert relerr(got[k], exp[k]).max() <= TOL


@pytest.mark.parametrize("line", [9.4, 32.1, 41.5])
def test_kr83m_yields(ctx, ref, line):
    """GetYieldKr83m (NEST.cpp:954-1044) at fixed decay separation: deterministic, 1e-12"""
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    F = np.array(FIELDS, float)
    E = np.full(F.size, line)
    for dT in (20.0, 50.0, 154.4, 400.0, 1000.0, 5000.0):
        mo.massNum, mo.atomNum = dT, dT
        for rho in (RHO_LUX, 2.7):
            got = ctx.yields(det, mo, capi.Kr83m, E, F, rho)
            exp = ref.yields(0, capi.Kr83m, E, F, rho, dT, dT, list(mo.NRYieldsParam), list(mo.ERYieldsParam)).T
            for k in range(6):
                assert relerr(got[k], exp[k]).max() <= TOL, (line, dT, rho, k, got[k], exp[k])
    # min > max is the reference's "ERROR: min greater than max" (:957-959), except on the 32.1 keV line
    mo.massNum, mo.atomNum = 100.0, 300.0
    if line != 32.1:
        with pytest.raises(capi.NestB200Error, match="min greater than max"):
            ctx.yields(det, mo, capi.Kr83m, E, F, RHO_LUX)


def test_survey_pinned_values(ctx):
    """values captured from the compiled reference in SURVEY.md 8(c)"""
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    mo.massNum = 131.293
    mo.er_model = 0
    b = ctx.yields(det, mo, capi.beta, [5.0, 0.5, 20.0, 100.0], [180.0, 180.0, 50.0, 1000.0], RHO_LUX)
    exp = [(205.88604645954106, 165.20770456738973, 0.050298182684348262),
           (0.64889386770322943, 36.460481234989849, 0.0051337477773812801),
           (1113.8418225328433, 370.53318157487979, 0.15339221999832758),
           (2806.3962674277627, 4615.4787531108532, 0.18202453502051572)]
    for i, e in enumerate(exp):
        for k in range(3):
            assert abs(b[k][i] - e[k]) <= TOL * abs(e[k])
    mo.er_model = -1
    n = ctx.yields(det, mo, capi.NR, [1.0, 10.0, 50.0, 0.3], [180.0, 180.0, 400.0, 180.0], RHO_LUX)
    expn = [(3.8175326853690859, 6.8511840288214803, 0.50923264205705066, 0.14374691954077559),
            (80.68553759012174, 57.878421270493256, 0.81936822679002697, 0.18669670194818133),
            (632.60433974166733, 181.5986274608284, 0.83580561966331435, 0.21940627266003407),
            (0.88939281415746496, 1.1501592836462304, 0.76402647275054425, 0.091600936068923555)]
    for i, e in enumerate(expn):
        for k in range(4):
            assert abs(n[k][i] - e[k]) <= TOL * abs(e[k])


@pytest.mark.parametrize("trial", [0, 1, 2])
def test_yields_with_user_parameter_vectors(ctx, ref, trial):
    """the mean-yield models with the parameter vectors a user fits instead of the defaults: ERYieldsParam entries >= 0
    replace the built-in field / density fits of the beta model one by one (NEST.cpp:1135-1180: each m_i has its own
    `< 0 means default` switch), NRYieldsParam moves every term of the NR model (:773-854) -- 1e-12 against the reference"""
    rng = np.random.default_rng(100 + trial)
    det = capi.detector("LUX_Run03")
    capi.prepare(det, 180.0)
    mo = capi.model(1)
    nr0 = np.array(list(mo.NRYieldsParam))
    nr = nr0 * (1.0 + 0.15 * rng.uniform(-1, 1, nr0.size))
    # plausible explicit values for m1..m10 of the beta model; a random subset stays at -1 (built-in fit)
    er = np.array([30.0, 75.0, 1.2, 3.0, 25.0, 0.0, 60.0, 4.0, 0.3, 0.04]) * (1.0 + 0.2 * rng.uniform(-1, 1, 10))
    keep_default = rng.uniform(size=10) < (0.0 if trial == 0 else 0.4)
    er[keep_default] = -1.0
    er[5] = 0.0 if not keep_default[5] else -1.0
    for i in range(12):
        mo.NRYieldsParam[i] = nr[i]
    for i in range(10):
        mo.ERYieldsParam[i] = er[i]
    mo.massNum, mo.atomNum = 131.293, 54.0
    for name, species, er_model, which_ref, mult, erange in [
            ("NR", capi.NR, -1, 0, 1.0, (0.05, 330.0)), ("beta", capi.beta, -1, 0, 1.0, (0.01, 3000.0)),
            ("betaGR_mult", capi.beta, 0, 1, 1.1, (0.01, 3000.0)), ("ERweighted", capi.beta, 2, 3, 1.0, (0.01, 3000.0))]:
        mo.er_model, mo.multFact = er_model, mult
        E, F = grid(erange[0], erange[1], 200, FIELDS)
        got = ctx.yields(det, mo, species, E, F, RHO_LUX)
        exp = ref.yields(which_ref, species, E, F, RHO_LUX, 131.293, 54.0, list(nr), list(er), mult).T
        for k, col in enumerate(["Nph", "Ne", "NexONi", "L", "field", "dT"]):
            r = relerr(got[k], exp[k])
            assert r.max() <= TOL, "%s %s trial %d: max rel err %.3e at E=%g F=%g (gpu %r ref %r) er=%s" % (
                name, col, trial, r.max(), E[r.argmax()], F[r.argmax()], got[k][r.argmax()], exp[k][r.argmax()], er)
        assert np.array_equal(got[0] == 0, exp[0] == 0) and np.array_equal(got[1] == 0, exp[1] == 0)
