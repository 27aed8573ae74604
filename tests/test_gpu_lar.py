"""Liquid-argon yields and fluctuations on the GPU (BASELINE config 5) against the compiled
reference: LArNEST::FullCalculation src/LArNEST.cpp:17-212, 363-406."""
import numpy as np
import pytest
from scipy import stats

from nest_b200 import capi

pytestmark = pytest.mark.gpu


def benchmark_grid():
    """examples/LArNEST/LArNESTMeanYieldsBenchmarks.cpp:31-41"""
    step = (1000 - .1) / 50000
    return .1 + step * np.arange(50000)


@pytest.mark.parametrize("species", [0, 1, 2])
def test_lar_mean_yields_1e12(ctx, ref, species):
    E = benchmark_grid()
    for F in (1.0, 100.0, 500.0, 1000.0):
        y, _ = ctx.lar(species, E, F, 1.393, want_fluct=False)
        ry, _ = ref.lar(1, species, E, F, 1.393)
        for k in range(9):
            a, b = y[k], ry[:, k]
            d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
            assert d.max() <= 1e-12, (species, F, k, d.max())


def test_lar_exact_lindhard_branch(ctx, ref):
    """`Lindhard == 1.` (LArNEST.cpp:374) is an exact comparison whose outcome for ER flips with
    the last bit of E: the GPU must take the same branch for every benchmark energy."""
    E = benchmark_grid()
    y, _ = ctx.lar(1, E, 500.0, 1.393, want_fluct=False)
    ry, _ = ref.lar(1, 1, E, 500.0, 1.393)
    assert np.array_equal(y[7] == 1.0, ry[:, 7] == 1.0)
    assert np.array_equal(y[7], ry[:, 7])                 # five IEEE operations: bit-identical
    frac = (y[7] == 1.0).mean()
    assert 0.5 < frac < 0.65                              # true for 57.7 % of the grid


@pytest.mark.parametrize("species,E0", [(0, 20.0), (1, 10.0), (1, 10.000000000000002), (1, 300.0), (2, 5000.0)])
def test_lar_fluctuations_distribution(ctx, ref, species, E0):
    n = 100000
    E = np.full(n, E0)
    _, f = ctx.lar(species, E, 500.0, 1.393, seed=3)
    _, rf = ref.lar(5, species, E, 500.0, 1.393)
    for k, name in enumerate(["Nph", "Ne", "Nex", "Nion"]):
        p = stats.ks_2samp(f[k], rf[:, k]).pvalue
        assert p > 1e-4, (species, E0, name, p, f[k].mean(), rf[:, k].mean())


def test_lar_batch_invariance(ctx):
    E = np.linspace(1.0, 500.0, 5000)
    _, a = ctx.lar(1, E, 500.0, 1.393, seed=9, first_event=0)
    _, b1 = ctx.lar(1, E[:1234], 500.0, 1.393, seed=9, first_event=0)
    _, b2 = ctx.lar(1, E[1234:], 500.0, 1.393, seed=9, first_event=1234)
    assert np.array_equal(a, np.hstack([b1, b2]))
    with pytest.raises(capi.NestB200Error):
        ctx.lar(3, E, 500.0, 1.393)      # dE/dx and LET models are outside this path


@pytest.mark.parametrize("species", [0, 1, 2])
def test_lar_edge_energies(ctx, ref, species):
    """the bit-built logarithm of k_lar only takes positive normal finite energies; zero, negative, subnormal, huge and
    non-finite ones go through libm and must give the reference's zeros / inf / NaN column by column"""
    E = np.array([0.0, -1.0, 1e-310, 2.3e-308, 1e-300, 1e-30, 1e30, 1e300, np.inf, np.nan, 0.1, 1000.0])
    if species == 1:
        # the reference's own FullCalculation does not return for ER energies whose `Lindhard == 1.` branch hands
        # std::binomial_distribution a count that does not fit an int64 (or a NaN): the tiny and the ordinary only
        E = np.array([1e-310, 2.3e-308, 1e-300, 1e-30, 0.1, 1000.0])
    with np.errstate(all="ignore"):
        y, _ = ctx.lar(species, E, 500.0, 1.393, want_fluct=False)
        ry, _ = ref.lar(1, species, E, 500.0, 1.393)
        for k in range(9):
            a, b = y[k], ry[:, k]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (species, k, a, b)
            assert np.array_equal(np.isinf(a), np.isinf(b)), (species, k, a, b)
            m = np.isfinite(a) & np.isfinite(b)
            d = np.abs(a[m] - b[m]) / np.maximum(np.maximum(np.abs(a[m]), np.abs(b[m])), 1e-300)
            assert d.size == 0 or d.max() <= 1e-12, (species, k, a, b)


def test_lar_queued_binomial_draws_equal_the_inline_sampler(ctx):
    """ER events on the `Lindhard == 1.` branch draw Nion ~ Binomial(Nq, 1 / 1.21).  k_lar queues those draws per warp
    and makes them one candidate per lane and pass (nb_lar.cuh); candidate k of an event's draw is Philox block k of
    its stream whoever makes it, so every Nion must be what the in-line sampler returns for that event and Nq --
    and the moments must be the binomial's"""
    E = np.linspace(0.5, 900.0, 40_000)
    y, f = ctx.lar(1, E, 500.0, 1.393, seed=21, first_event=1000)
    exact = y[7] == 1.0
    assert 0.4 < exact.mean() < 0.7
    nion, nex = f[3], f[2]
    nq = (nion + nex)[exact].astype(np.int64)
    alf = 1.0 / (1.0 + 0.21)
    # the in-line sampler on the same events (draw i of the hook = event first_event + i)
    idx = np.nonzero(exact)[0]
    n_all = np.zeros(E.size, np.int64)
    n_all[idx] = nq
    inline = ctx.debug_binomial_list(n_all, alf, 1, seed=21, first_event=1000)
    assert np.array_equal(inline[idx], nion[exact].astype(np.int64))
    # batch split anywhere (the queue's contents differ, the results must not)
    _, f2 = ctx.lar(1, E[777:], 500.0, 1.393, seed=21, first_event=1777)
    assert np.array_equal(f[:, 777:], f2)
    # moments of Nion given Nq
    z = (nion[exact] - nq * alf) / np.sqrt(np.maximum(nq, 1) * alf * (1 - alf))
    big = nq > 50
    assert abs(z[big].mean()) < 4.0 / np.sqrt(big.sum())
    assert abs(z[big].var() - 1.0) < 0.03


@pytest.mark.parametrize("species", [0, 1, 2])
def test_lar_field_and_density_corners(ctx, ref, species):
    """fields from zero to far beyond the benchmarks' 1 kV/cm and other densities: the field-only factors kept per thread
    (LArFieldTerms) are re-evaluated at every change and must give the reference's values, infinities and NaN alike"""
    E = np.array([0.5, 3.0, 20.0, 150.0, 900.0])
    for rho in (1.393, 1.2, 1.5):
        for F in (0.0, 1e-3, 0.3, 7.0, 2.5e3, 1e5):
            # one call with the field changing from event to event (the memo's worst case), in both directions
            Fs = np.array([F, 1.0, F, 1000.0, F])
            with np.errstate(all="ignore"):
                y, _ = ctx.lar(species, E, Fs, rho, want_fluct=False)
                ry = np.vstack([ref.lar(1, species, E[i:i + 1], float(Fs[i]), rho)[0] for i in range(E.size)])
            for k in range(9):
                a, b = y[k], ry[:, k]
                assert np.array_equal(np.isnan(a), np.isnan(b)), (species, rho, F, k, a, b)
                assert np.array_equal(np.isinf(a), np.isinf(b)), (species, rho, F, k, a, b)
                m = np.isfinite(a) & np.isfinite(b)
                d = np.abs(a[m] - b[m]) / np.maximum(np.maximum(np.abs(a[m]), np.abs(b[m])), 1e-300)
                assert d.size == 0 or d.max() <= 1e-12, (species, rho, F, k, a, b)
