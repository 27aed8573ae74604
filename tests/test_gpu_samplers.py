"""The binomial sampler of the chain (device replacement of RandomGen::binom_draw,
src/RandomGen.cpp:132-134 = std::binomial_distribution) against the exact pmf.

Every regime of the sampler is covered: CDF inversion (n p < 10), the recursive f(k)/f(m)
branch of BTRD (small n p q), the squeeze / exact log-pmf branch (large n p q), p > 1/2
(reflected), and the sizes the chain really draws (S2 photons: n ~ 1e5..1e6)."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

CASES = [
    (5, 0.3), (40, 0.2), (1000, 0.005),            # inversion
    (60, 0.2), (100, 0.5), (200, 0.07), (150, 0.9),  # BTRD, |k-m| <= 15 almost always
    (2000, 0.12), (5000, 0.47), (30000, 0.83),      # BTRD, squeeze
    (250000, 0.094), (1200000, 0.173),              # S2-sized draws
]


@pytest.mark.parametrize("n,p", CASES)
def test_binomial_matches_pmf(ctx, n, p):
    count = 4_000_000
    k = ctx.debug_binomial(n, p, count, seed=12345 + n)
    assert k.min() >= 0 and k.max() <= n
    lo, hi = int(k.min()), int(k.max())
    obs = np.bincount(k - lo, minlength=hi - lo + 1).astype(float)
    ks = np.arange(lo, hi + 1)
    exp = stats.binom.pmf(ks, n, p) * count
    # pool the tails so that every cell expects >= 20 draws
    keep = exp >= 20
    i0, i1 = np.argmax(keep), len(keep) - np.argmax(keep[::-1])
    o = np.concatenate([[obs[:i0].sum()], obs[i0:i1], [obs[i1:].sum()]])
    e = np.concatenate([[stats.binom.cdf(lo + i0 - 1, n, p) * count], exp[i0:i1],
                        [stats.binom.sf(lo + i1 - 1, n, p) * count]])
    m = e > 0
    chi2 = ((o[m] - e[m]) ** 2 / e[m]).sum()
    dof = m.sum() - 1
    pval = stats.chi2.sf(chi2, dof)
    # moments to ~5 sigma of their sampling error
    mu, var = n * p, n * p * (1 - p)
    assert abs(k.mean() - mu) < 5 * np.sqrt(var / count)
    assert abs(k.var() - var) < 5 * var * np.sqrt(2.0 / count) + 5 * np.sqrt(var / count)
    assert pval > 1e-4, (n, p, chi2, dof, pval)


def test_binomial_edges(ctx):
    assert (ctx.debug_binomial(0, 0.3, 1000) == 0).all()
    assert (ctx.debug_binomial(100, 0.0, 1000) == 0).all()
    assert (ctx.debug_binomial(100, 1.0, 1000) == 100).all()
    assert (ctx.debug_binomial(-5, 0.5, 1000) == 0).all()
    # disjoint event ranges give disjoint streams; the same range reproduces
    a = ctx.debug_binomial(5000, 0.3, 10000, seed=7, first_event=0)
    b = ctx.debug_binomial(5000, 0.3, 10000, seed=7, first_event=0)
    c = ctx.debug_binomial(5000, 0.3, 5000, seed=7, first_event=5000)
    assert (a == b).all() and (a[5000:] == c).all()


def test_hit_loop_normal_is_standard_normal(ctx):
    """The S1 hit loop's ziggurat normal (core + wedge + tail, drawn exactly as k_hits and its slow
    finisher draw it) against N(0,1): KS, chi2 over 200 equiprobable cells, moments to 5 sigma,
    the mass beyond the ziggurat's R (tail algorithm) and the wedge/tail share."""
    slots = 7_000_000
    z, kind = ctx.debug_normal(slots, seed=2024)
    n = z.size
    assert np.isfinite(z).all()
    # the rare completions are 0.23 % of the draws (1 - mean of x_{i+1}/x_i over the 2048 layers)
    assert abs(kind.mean() - 0.0022711) < 5 * np.sqrt(0.0022711 / n)
    assert stats.kstest(z[:4_000_000], "norm").pvalue > 1e-4
    edges = stats.norm.ppf(np.linspace(0, 1, 201)[1:-1])
    obs = np.bincount(np.searchsorted(edges, z), minlength=200).astype(float)
    chi2 = ((obs - n / 200) ** 2 / (n / 200)).sum()
    assert stats.chi2.sf(chi2, 199) > 1e-4, chi2
    # the wedge / tail completions on their own: |z| of those draws against the exact law of a completion
    # is covered by the cells above only through their 0.4 % weight, so test where they dominate --
    # the outermost cells, where every draw beyond x_2 of the ziggurat is a completion or a base-strip core
    # the completions alone, weighted back in, must not distort anything: chi2 of |z| in the tail region
    R = 4.216370409511897
    for cut in (3.0, 3.6, R, 4.5):
        exp = 2 * stats.norm.sf(cut) * n
        got = (np.abs(z) > cut).sum()
        assert abs(got - exp) < 5 * np.sqrt(exp) + 1, (cut, got, exp)
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs(stats.skew(z)) < 5 * np.sqrt(6 / n)
    assert abs(stats.kurtosis(z)) < 5 * np.sqrt(24 / n)
    assert abs((z ** 6).mean() - 15) < 5 * np.sqrt((10395 - 225) / n)
    # the three normals of a slot (one Philox block), and neighbouring slots, are uncorrelated
    a, b, c = z[0::3], z[1::3], z[2::3]
    for u, v in ((a, b), (b, c), (a, c), (c[:-1], a[1:])):
        assert abs(np.corrcoef(u, v)[0, 1]) < 5 / np.sqrt(slots)
    # reproducible, and a function of (seed, event, slot) only
    z2, _ = ctx.debug_normal(2048, seed=2024, first_event=1)
    assert (z2 == z[3 * 1024:3 * 1024 + 3 * 2048]).all()


@pytest.mark.parametrize("n,p", [(1, 0.173), (7, 0.173), (64, 0.173), (79, 0.173), (350, 0.2), (2047, 0.05), (1000, 0.9),
                                  (4096, 0.173), (5000, 0.173), (65535, 0.173)])
def test_fixed_p_binomial_matches_pmf(ctx, n, p):
    """The alias-table sampler of the per-event double-phe count against the exact pmf."""
    count = 4_000_000
    k = ctx.debug_binomial_fixed(n, p, count, seed=99 + n)
    assert k.min() >= 0 and k.max() <= n
    obs = np.bincount(k, minlength=n + 1).astype(float)
    exp = stats.binom.pmf(np.arange(n + 1), n, p) * count
    keep = exp >= 20
    o = np.concatenate([obs[keep], [obs[~keep].sum()]])
    e = np.concatenate([exp[keep], [exp[~keep].sum()]])
    m = e > 0
    chi2 = ((o[m] - e[m]) ** 2 / e[m]).sum()
    assert stats.chi2.sf(chi2, max(m.sum() - 1, 1)) > 1e-4, (n, p, chi2)
    mu, var = n * p, n * p * (1 - p)
    assert abs(k.mean() - mu) < 5 * np.sqrt(var / count)
    assert (ctx.debug_binomial_fixed(0, p, 100) == 0).all()
