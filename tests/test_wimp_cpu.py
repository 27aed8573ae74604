"""WIMP source, host side (BASELINE config 4): nb200_wimp_drate / nb200_wimp_prep_iso against the compiled reference's
TestSpectra::WIMP_dRate (src/TestSpectra.cpp:251-390) and WIMP_prep_spectrum (:392-454).

The reference draws a random Xe isotope inside every WIMP_dRate call (:278), so its rate and its table are themselves
random; both are compared for the reference's own isotope draws (read back from its generator under the same seed).
No device needed."""
import ctypes as C

import numpy as np
import pytest

from nest_b200 import capi

N_INTEGRAL = 1_000_000  # terms of WIMP_prep_spectrum's integral (:420-422), one isotope draw each


def ours_drate(ER, mass, day, A):
    ok = C.c_int()
    out = np.empty(len(ER))
    for i, (x, a) in enumerate(zip(ER, A)):
        out[i] = capi.lib().nb200_wimp_drate(float(x), mass, day, int(a), C.byref(ok))
        assert ok.value == 1
    return out


@pytest.mark.parametrize("mass,day", [(50.0, 0.0), (5.0, 0.0), (8.0, 152.5), (500.0, 0.0), (50.0, 300.0)])
def test_wimp_drate_matches_reference(built, ref, mass, day):
    ER = np.concatenate([[0.0], np.exp(np.linspace(np.log(1e-3), np.log(150.0), 300))])
    v, A = ref.wimp_drate(1000 + int(mass), ER, mass, day)
    assert set(A) <= {124, 126, 128, 129, 130, 131, 132, 134, 136} and len(set(A)) >= 4
    mine = ours_drate(ER, mass, day, A)
    nz = v != 0
    assert np.array_equal(mine == 0, v == 0)
    assert np.max(np.abs(mine[nz] - v[nz]) / np.abs(v[nz])) <= 1e-12
    assert v[0] > v[nz][-1] > 0 or not nz[-1]


@pytest.mark.parametrize("mass,eStep,seed", [(50.0, 5.0, 5), (50.0, 5.0, 11), (8.0, 5.0, 3), (200.0, 5.0, 2), (30.0, 2.0, 2)])
def test_wimp_table_matches_reference_for_its_isotope_draws(built, ref, mass, eStep, seed):
    b, e, integ, xmax, div = ref.wimp_prep(seed, mass, eStep, 0.0)
    iso = ref.sampler(6, seed, N_INTEGRAL + 1100).astype(np.int32)  # the reference's own SelectRanXeAtom sequence
    w = capi.WimpTable(mass, eStep, 0.0, isotopes=iso)
    assert w.divisor == div and w.xMax == xmax
    for mine, theirs in ((w.base, b), (w.exponent, e)):
        # the entry at which the reference stops (base or exponent not positive and finite, :433-449) may be NaN / inf
        assert np.array_equal(np.isnan(mine), np.isnan(theirs)) and np.array_equal(np.isinf(mine), np.isinf(theirs))
        nz = np.isfinite(theirs) & (theirs != 0)
        assert np.array_equal(mine == 0, theirs == 0)
        assert np.max(np.abs(mine[nz] - theirs[nz]) / np.abs(theirs[nz])) <= 1e-12
    assert abs(w.integral - integ) <= 1e-12 * integ
    # a different isotope sequence gives a (slightly) different table: the comparison above is not vacuous
    w2 = capi.WimpTable(mass, eStep, 0.0, seed=seed)
    assert not np.array_equal(w2.base, w.base)
    assert abs(w2.integral / w.integral - 1) < 5e-3   # ... but the same rate integral to the isotope-mixture noise


@pytest.mark.parametrize("mass,eStep", [(200.0, 3.0), (8.0, 0.5)])
def test_wimp_prep_fails_where_the_reference_throws(built, ref, mass, eStep):
    """"ERROR: WIMP E_step is too small (or large)!" (TestSpectra.cpp:436-441) -> NB200_EPHYSICS"""
    with pytest.raises(RuntimeError):
        ref.wimp_prep(2, mass, eStep, 0.0)
    with pytest.raises(capi.NestB200Error) as ei:
        capi.WimpTable(mass, eStep, 0.0, seed=2)
    assert ei.value.code == capi.EPHYSICS


def test_wimp_prep_iso_rejects_a_short_list(built):
    with pytest.raises(capi.NestB200Error) as ei:
        capi.WimpTable(50.0, 5.0, 0.0, isotopes=np.full(10, 131, np.int32))
    assert ei.value.code == capi.EINVAL


def test_poisson_event_count_is_poisson(built, ref):
    """execNEST.cpp:482-488: numEvts = poisson_draw(integral * exposure * xsec / 1e-36); ours draws from the Philox
    host stream of the seed.  Both against the exact pmf, and against each other."""
    from scipy import stats
    for mean in (3.7, 78.0, 30397.7):
        n = 20000
        mine = np.array([capi.lib().nb200_poisson(mean, s) for s in range(1, n + 1)], dtype=float)
        theirs = ref.poisson(99, n, mean)
        for x in (mine, theirs):
            assert abs(x.mean() - mean) < 5 * np.sqrt(mean / n)
            assert abs(x.var() - mean) < 6 * mean * np.sqrt(2.0 / n) + 6 * np.sqrt(mean / n)
        assert stats.ks_2samp(mine, theirs).pvalue > 1e-4
        lo, hi = int(stats.poisson.ppf(1e-3, mean)), int(stats.poisson.ppf(1 - 1e-3, mean))
        edges = np.unique(np.linspace(lo, hi, 25).astype(int))
        obs = np.histogram(mine, bins=np.concatenate([[-0.5], edges + 0.5, [1e18]]))[0]
        cdf = np.concatenate([[0.0], stats.poisson.cdf(edges, mean), [1.0]])
        exp = np.diff(cdf) * n
        m = exp > 5
        chi2 = ((obs[m] - exp[m]) ** 2 / exp[m]).sum()
        assert stats.chi2.sf(chi2, m.sum() - 1) > 1e-4, (mean, chi2)
