"""Band post-processing on the host (nb200_band_chi2, nb200_band_leakage) without a GPU: the two
uses rootNEST makes of a band that need no ROOT fit, checked against numpy restatements of
reference examples/rootNEST.cpp:600-643 (goodness of fit) and :880-915 (leakage by counting)."""
import numpy as np
import pytest

from nest_b200 import capi, shard


def synthetic_signals(seed, n, mu0, slope, sigma):
    """(S1, S2) pairs whose log10(S2/S1) is N(mu0 + slope*S1, sigma) -- a band-shaped cloud"""
    rng = np.random.default_rng(seed)
    s1 = rng.uniform(1.5, 60.0, n)
    y = rng.normal(mu0 + slope * s1, sigma)
    return s1, s1 * 10.0 ** y


def band_of(ana, s1, s2):
    return shard.band_from_words(ana, shard.band_words_from_signals(ana, 2, s1, s2))


def chi2_restated(band, band2, useS2, free_param):
    """examples/rootNEST.cpp:619-643 with the error columns this library's bands carry:
    [4] = sigma/sqrt(N) for the mean, [4]/sqrt(2) for the width"""
    nb = len(band)
    chi = np.zeros(4)
    for i in range(nb):
        if not (band[i][4] > 0 and band2[i][4] > 0):
            continue
        if (useS2 == 0 and abs(band[i][2]) < 5.) or useS2 != 0:
            err = np.sqrt(band[i][4] ** 2 + band2[i][4] ** 2)
            chi[0] += ((band2[i][2] - band[i][2]) / err) ** 2
            err = np.sqrt(band[i][4] ** 2 / 2 + band2[i][4] ** 2 / 2)
            chi[1] += ((band2[i][3] - band[i][3]) / err) ** 2
            chi[2] += 100. * (band2[i][2] - band[i][2]) / band2[i][2]
            chi[3] += 100. * (band2[i][3] - band[i][3]) / band2[i][3]
    dof = nb - abs(free_param)
    chi[0] /= dof - 1
    chi[1] /= dof - 1
    chi[2] /= nb
    chi[3] /= nb
    return np.array([chi[0], chi[1], 0.5 * (chi[0] + chi[1]), np.sqrt(chi[0] * chi[1]), -chi[2], -chi[3]])


def test_band_chi2_follows_rootnest(built):
    ana = capi.analysis(0)
    a = band_of(ana, *synthetic_signals(1, 200000, 2.6, -0.006, 0.11)).finalize()
    b = band_of(ana, *synthetic_signals(2, 150000, 2.6, -0.006, 0.11)).finalize()
    c = band_of(ana, *synthetic_signals(3, 150000, 2.63, -0.006, 0.13)).finalize()
    # bins above S1 = 60 are empty on both sides (N = 0: NaN statistics) and must be skipped, not propagated
    assert np.isnan(a[-1, 2]) and np.isfinite(a[10, 2])
    for other in (b, c):
        got = capi.band_chi2(a, other, useS2=0, free_param=2)
        np.testing.assert_allclose(got, chi2_restated(a, other, 0, 2), rtol=1e-12, atol=0)
    same = capi.band_chi2(a, b)
    diff = capi.band_chi2(a, c)
    assert same[0] < 2.0 and same[1] < 2.0, same        # two samples of one band agree
    assert diff[0] > 20.0 and diff[1] > 20.0, diff      # a shifted, wider band does not
    assert diff[4] < 0 and diff[5] < 0                  # a - b < 0 in both mean and width
    # the reference's own argument order: band2 is the one compared against
    swapped = capi.band_chi2(c, a)
    assert swapped[0] == pytest.approx(diff[0], rel=1e-12) and swapped[4] > 0
    # free parameters only change the divisor
    fp5 = capi.band_chi2(a, c, free_param=-5)
    assert fp5[0] == pytest.approx(diff[0] * (ana.numBins - 3) / (ana.numBins - 6), rel=1e-12)


def test_band_chi2_outlier_and_binning_rules(built):
    ana = capi.analysis(0)
    a = band_of(ana, *synthetic_signals(4, 50000, 2.6, -0.006, 0.11)).finalize()
    b = band_of(ana, *synthetic_signals(5, 50000, 2.6, -0.006, 0.11)).finalize()
    base = capi.band_chi2(a, b)
    # useS2 == 0 ignores bins whose simulated mean is >= 5 in magnitude (rootNEST.cpp:629); other modes do not
    a2 = a.copy()
    a2[7, 2] = 6.0
    assert capi.band_chi2(a2, b, useS2=0)[0] < base[0] + 1e-9
    assert capi.band_chi2(a2, b, useS2=1)[0] > 100.0
    # bin centres further apart than 0.05: the reference's "Binning doesn't match for GoF calculation", return 1
    b2 = b.copy()
    b2[:, 0] += 0.06
    with pytest.raises(capi.NestB200Error) as e:
        capi.band_chi2(a, b2)
    assert e.value.code == capi.EINVAL
    b2[:, 0] = b[:, 0] + 0.04
    assert capi.band_chi2(a, b2)[0] == pytest.approx(base[0], rel=1e-12)


def leakage_restated(ana, s1, s2, centroid):
    """examples/rootNEST.cpp:880-915 with NRbandCenter = 1: per S1 bin, the events GetBand kept
    (execNEST.cpp:1540-1570; out-of-range y enter as 0) counted below that bin's centroid"""
    nb = ana.numBins
    w = (ana.maxS1 - ana.minS1) / nb
    below, size = np.zeros(nb), np.zeros(nb)
    for i in range(nb):
        lo = ana.minS1 + i * w
        m = (s1 > lo) & (s1 <= lo + w) & (s2 >= 0)
        y = np.log10(s2[m] / s1[m])
        y = np.where((y > ana.logMin) & (y < ana.logMax), y, 0.0)
        below[i] = (y < centroid[i]).sum()
        size[i] = m.sum()
    return below, size


def test_band_leakage_counts_like_rootnest(built):
    ana = capi.analysis(0)
    assert ana.logBins == 30
    ana.logBins = 300                                    # 0.01 in log10(S2/S1): the caller chooses the resolution
    res = (ana.logMax - ana.logMin) / ana.logBins
    er1, er2 = synthetic_signals(6, 400000, 2.6, -0.006, 0.11)
    nr1, nr2 = synthetic_signals(7, 100000, 2.25, -0.002, 0.12)
    er, nr = band_of(ana, er1, er2), band_of(ana, nr1, nr2)
    centroid = nr.finalize()[:, 2]                      # the NR band's mean per S1 bin
    filled = er.accepted > 0
    cen = np.where(np.isfinite(centroid), centroid, 0.0)
    leak, hi, lo, tot = er.leakage(cen)
    below, size = leakage_restated(ana, er1, er2, cen)
    np.testing.assert_array_equal(size, er.accepted.astype(float))
    N = size[filled]
    # the histogram resolves the centroid to one log bin: the count may differ by the content of the bin that
    # holds it, never by more
    rows = er.hist2d.reshape(ana.numBins, ana.logBins).astype(float)
    jc = np.clip(((cen - ana.logMin) / res).astype(int), 0, ana.logBins - 1)
    slack = rows[np.arange(ana.numBins), jc]
    assert np.all(np.abs(leak[filled] * N - below[filled]) <= slack[filled] + 1e-9)
    assert np.abs(leak[filled] * N - below[filled]).max() > 0 or slack[filled].max() == 0
    # overall: some 10^-3 of this ER cloud leaks, and the histogram estimate agrees within a few per cent
    assert tot[1] == size.sum() and tot[0] == pytest.approx((leak[filled] * N).sum(), rel=1e-12)
    assert tot[2] == pytest.approx(below.sum() / size.sum(), rel=0.08)
    assert 1e-4 < tot[2] < 0.05
    # Poisson error bars exactly as rootNEST.cpp:898-903
    k = leak[filled] * N
    np.testing.assert_allclose(hi[filled], (k + np.sqrt(k)) / (N - np.sqrt(N)) - leak[filled], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(lo[filled], leak[filled] - (k - np.sqrt(k)) / (N + np.sqrt(N)), rtol=1e-12, atol=1e-15)


def test_band_leakage_edges(built):
    ana = capi.analysis(0)
    s1, s2 = synthetic_signals(8, 50000, 2.6, -0.006, 0.11)
    # a tenth of the events far outside (logMin, logMax): GetBand pushes y = 0 for them (execNEST.cpp:1551-1555)
    s2[::10] = s1[::10] * 10.0 ** 4.5
    b = band_of(ana, s1, s2)
    nb = ana.numBins
    filled = b.accepted > 0
    N = b.accepted[filled].astype(float)
    in_hist = b.hist2d.reshape(nb, ana.logBins).sum(axis=1)[filled].astype(float)
    assert np.all(in_hist < N)
    # a centroid above everything: all events are below it, the y = 0 ones included
    leak, _, _, tot = b.leakage(np.full(nb, ana.logMax))
    np.testing.assert_allclose(leak[filled], 1.0, rtol=1e-12)
    assert tot[0] == tot[1] == N.sum()
    # a centroid of 0: nothing inside the histogram is below it, and y = 0 is not below 0 (`<`, rootNEST.cpp:893)
    leak, _, _, tot = b.leakage(np.zeros(nb))
    assert np.all(leak[filled] == 0.0) and tot[0] == 0.0
    # a centroid at the lower edge of the log axis (0.6 > 0): only the y = 0 events are below it
    leak, _, _, _ = b.leakage(np.full(nb, ana.logMin))
    np.testing.assert_allclose(leak[filled] * N, N - in_hist, rtol=0, atol=1e-3)
    # the centroid inside a log bin splits that bin linearly: monotone in the centroid
    c = np.linspace(2.0, 2.8, 33)
    tots = [b.leakage(np.full(nb, x))[3][2] for x in c]
    assert np.all(np.diff(tots) >= 0) and tots[0] < 0.2 < 0.8 < tots[-1]
    # bad arguments
    no_hist = capi.Band(ana)
    no_hist.c.logBins = 0
    with pytest.raises(capi.NestB200Error):
        no_hist.leakage(np.zeros(nb))


def test_band_gaussian_fit_matches_least_squares(built):
    """nb200_band_fit_gauss against scipy's Levenberg-Marquardt on the same histogram rows with the same
    weights (1/content, empty bins left out, function at the bin centres -- TH1::Fit's default chi^2)"""
    from scipy.optimize import curve_fit
    ana = capi.analysis(0)
    ana.logBins = 150
    b = band_of(ana, *synthetic_signals(9, 600000, 2.6, -0.006, 0.11))
    fit = b.fit_gauss()
    mom = b.finalize()
    nb, lb = ana.numBins, ana.logBins
    w = (ana.logMax - ana.logMin) / lb
    x = ana.logMin + (np.arange(lb) + 0.5) * w
    rows = b.hist2d.reshape(nb, lb).astype(float)
    f = lambda x, A, m, s: A * np.exp(-0.5 * ((x - m) / s) ** 2)
    checked = 0
    for j in range(nb):
        if b.accepted[j] < 2000:
            assert fit[j, 7] == 0 or fit[j, 2] > 0
            continue
        k = rows[j] > 0
        popt, pcov = curve_fit(f, x[k], rows[j][k], p0=(rows[j].max(), mom[j, 2], mom[j, 3]), sigma=np.sqrt(rows[j][k]),
                               absolute_sigma=True, xtol=1e-14, ftol=1e-14, gtol=1e-14)
        np.testing.assert_allclose(fit[j, :3], popt, rtol=2e-6)
        np.testing.assert_allclose(fit[j, 3:5], np.sqrt(np.diag(pcov))[1:], rtol=1e-4)
        chi = (((rows[j][k] - f(x[k], *popt)) ** 2) / rows[j][k]).sum() / (k.sum() - 3)
        assert fit[j, 5] == pytest.approx(chi, rel=1e-6) and fit[j, 7] == k.sum()
        # the sample is Gaussian: fit and moments tell the same story, the fit is good, and the errors are the
        # size of the moments' own
        assert abs(fit[j, 1] - mom[j, 2]) < 5 * mom[j, 4] and abs(fit[j, 2] - mom[j, 3]) < 5 * mom[j, 4]
        assert 0.2 < fit[j, 5] < 3.0
        assert 0.5 * mom[j, 4] < fit[j, 3] < 2.0 * mom[j, 4]
        # the reference's own figure (rootNEST.cpp:1346-1357), restated with its indexing
        g = 0.0
        for kk in range(lb):
            model = fit[j, 0] * np.exp(-0.5 * (ana.logMin + kk * w - fit[j, 1]) ** 2 / fit[j, 2] ** 2)
            content = float(b.accepted[j]) - rows[j].sum() if kk == 0 else rows[j][kk - 1]
            g += (content - model) ** 2 / max(np.float32(model + content), np.float32(1.0))
        assert fit[j, 6] == pytest.approx(g / (lb - 4.0), rel=1e-9)
        checked += 1
    assert checked > 40
    # bins without entries: a row of zeros, no NaN anywhere
    assert np.all(np.isfinite(fit)) and np.all(fit[b.accepted == 0] == 0)
    # too coarse an axis is refused
    coarse = capi.analysis(0)
    coarse.logBins = 3
    with pytest.raises(capi.NestB200Error):
        band_of(coarse, *synthetic_signals(9, 1000, 2.6, -0.006, 0.11)).fit_gauss()


def test_band_gaussian_leakage_follows_rootnest(built):
    from scipy.special import erf
    ana = capi.analysis(0)
    ana.logBins = 150
    er = band_of(ana, *synthetic_signals(10, 600000, 2.6, -0.006, 0.11)).fit_gauss()
    nr = band_of(ana, *synthetic_signals(11, 300000, 2.25, -0.002, 0.12)).fit_gauss()
    out = capi.band_leakage_gauss(er, nr)
    ok = (er[:, 7] > 0) & (nr[:, 7] > 0)
    assert ok.sum() > 40
    e, r = er[ok], nr[ok]
    # rootNEST.cpp:717-736 with band = ER (mean [2], sigma [3], errors [5], [6]) and band2 = NR
    n_sig = (e[:, 1] - r[:, 1]) / e[:, 2]
    up = (e[:, 1] - e[:, 3] - r[:, 1] - r[:, 3]) / (e[:, 2] + e[:, 4])
    dn = (e[:, 1] + e[:, 3] - r[:, 1] + r[:, 3]) / (e[:, 2] - e[:, 4])
    leak = (1 - erf(n_sig / np.sqrt(2))) / 2
    np.testing.assert_allclose(out[ok, 0], n_sig, rtol=1e-13)
    np.testing.assert_allclose(out[ok, 1], leak, rtol=1e-12)
    np.testing.assert_allclose(out[ok, 2], np.abs((1 - erf(up / np.sqrt(2))) / 2 - leak), rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(out[ok, 3], np.abs(leak - (1 - erf(dn / np.sqrt(2))) / 2), rtol=1e-9, atol=1e-18)
    # the synthetic bands are (0.35 - 0.004 S1) / 0.11 sigmas apart: the Gaussian leakage says so
    s1 = ana.minS1 + (np.arange(ana.numBins) + 0.5) * (ana.maxS1 - ana.minS1) / ana.numBins
    np.testing.assert_allclose(out[ok, 0], ((0.35 - 0.004 * s1) / 0.11)[ok], atol=0.08)
    # NUMSIGABV moves the NR line by that many NR sigmas
    shifted = capi.band_leakage_gauss(er, nr, num_sig_above=-1.0)
    np.testing.assert_allclose(shifted[ok, 0], (e[:, 1] - (r[:, 1] - r[:, 2])) / e[:, 2], rtol=1e-13)
    with pytest.raises(ValueError):
        capi.band_leakage_gauss(er[:, :7], nr[:, :7])


def test_skew_gaussian_fit_matches_least_squares(built):
    """nb200_hist_fit model 1 (rootNEST's "skewband", examples/rootNEST.cpp:1361-1364) against scipy's least squares
    with the same weights on the same histogram, and back to the parameters the sample was drawn with"""
    from scipy.optimize import curve_fit
    from scipy.special import erf
    from scipy.stats import skewnorm
    f = lambda x, A, xi, w, al: A / (w * np.sqrt(2 * np.pi)) * np.exp(-0.5 * ((x - xi) / w) ** 2) * (1 + erf(al * (x - xi) / (w * np.sqrt(2))))
    for seed, (xi, om, al, n) in enumerate([(2.45, 0.16, 1.8, 200000), (2.3, 0.12, -1.2, 80000), (2.0, 0.1, 1.0, 150000)]):
        y = skewnorm.rvs(al, loc=xi, scale=om, size=n, random_state=seed)
        mu, sd = y.mean(), y.std()
        lo, hi, bins = mu - 3 * sd, mu + 3 * sd, 60                      # rootNEST's per-slice axis (:1319-1323)
        h, edges = np.histogram(y, bins=bins, range=(lo, hi))
        x = 0.5 * (edges[1:] + edges[:-1])
        start = (n * (hi - lo) / bins, mu - 0.5 * sd, 1.2 * sd, 0.5)
        par, err, chi = capi.hist_fit(1, h, lo, hi, start)
        k = h > 0
        popt, pcov = curve_fit(f, x[k], h[k].astype(float), p0=start, sigma=np.sqrt(h[k]), absolute_sigma=True,
                               xtol=1e-14, ftol=1e-14, gtol=1e-14)
        np.testing.assert_allclose(par, popt, rtol=5e-5, atol=1e-6)
        np.testing.assert_allclose(err, np.sqrt(np.diag(pcov)), rtol=2e-3)
        c = (((h[k] - f(x[k], *par)) ** 2) / h[k]).sum() / (k.sum() - 4)
        assert chi == pytest.approx(c, rel=1e-9) and 0.3 < chi < 2.5
        # the sample's own parameters, within the fit's errors (bin-centre evaluation biases omega by ~w^2/24 omega)
        assert abs(par[1] - xi) < 6 * err[1] + 2e-3 and abs(par[2] - om) < 6 * err[2] + 2e-3 and abs(par[3] - al) < 6 * err[3] + 0.05
        assert par[0] == pytest.approx(h.sum() * (hi - lo) / bins, rel=0.02)   # the function integrates to A: N * bin width
    # model 0 through the same entry point equals the band fit's rows
    ana = capi.analysis(0)
    ana.logBins = 100
    b = band_of(ana, *synthetic_signals(12, 200000, 2.6, -0.006, 0.11))
    fit, mom = b.fit_gauss(), b.finalize()
    row = b.hist2d.reshape(ana.numBins, ana.logBins)[10]
    par, err, chi = capi.hist_fit(0, row, ana.logMin, ana.logMax, (row.max(), mom[10, 2], mom[10, 3]))
    np.testing.assert_allclose(par, fit[10, :3], rtol=1e-7)
    np.testing.assert_allclose(err[1:], fit[10, 3:5], rtol=1e-6)
    assert chi == pytest.approx(fit[10, 5], rel=1e-7)
    # too few filled bins: an error code, not a made-up answer
    with pytest.raises(capi.NestB200Error) as e:
        capi.hist_fit(1, np.array([0, 5, 9, 4, 0, 0], np.uint64), 0.0, 1.0, (10.0, 0.4, 0.2, 0.1))
    assert e.value.code == capi.EFIT
    with pytest.raises(ValueError):
        capi.hist_fit(0, row, 0.0, 1.0, (1.0, 2.0))


def test_skew_leakage_is_the_skew_normal_cdf(built):
    from scipy.stats import skewnorm
    rng = np.random.default_rng(5)
    n = 40
    xi, om, al = rng.uniform(2.2, 2.7, n), rng.uniform(0.08, 0.2, n), rng.uniform(-2.5, 2.5, n)
    om_err = om * 0.03
    line = xi - rng.uniform(0.5, 3.0, n) * om
    leak, hi, lo = capi.band_leakage_skew(xi, om, om_err, al, line)
    # the reference integrates Owen's T by a 2500-step midpoint rule (rootNEST.cpp:1587-1602): exact to ~1e-8
    np.testing.assert_allclose(leak, skewnorm.cdf(line, al, loc=xi, scale=om), atol=2e-7)
    np.testing.assert_allclose(hi, skewnorm.cdf(line, al, loc=xi, scale=om + om_err), atol=2e-7)
    np.testing.assert_allclose(lo, skewnorm.cdf(line, al, loc=xi, scale=om - om_err), atol=2e-7)
    assert np.all(hi >= leak - 1e-12) and np.all(lo <= leak + 1e-12)      # the line is below xi: wider = more leakage
    # alpha = 0 is the Gaussian leakage
    g, _, _ = capi.band_leakage_skew(xi, om, om_err, np.zeros(n), line)
    from scipy.special import erf
    np.testing.assert_allclose(g, 0.5 * (1 + erf((line - xi) / om / np.sqrt(2))), atol=1e-15)


def test_graph_fit_matches_least_squares(built):
    """nb200_graph_fit (TGraph::Fit's unweighted least squares for rootNEST's NR-centroid curves, rootNEST.cpp:855-878)
    against scipy from the reference's own start values"""
    from scipy.optimize import curve_fit
    rng = np.random.default_rng(3)
    x = 2.0 + np.arange(98.0)
    woods = lambda x, a, b, c, d: a / (x + b) + c * x + d
    y = woods(x, 8.0, 12.0, -2.0e-3, 1.9) + rng.normal(0, 0.004, x.size)
    par, chi = capi.graph_fit(2, x, y, (10., 15., -1.5e-3, 2.))              # SetParameters(10., 15., -1.5e-3, 2.), :858
    popt, _ = curve_fit(woods, x, y, p0=(10., 15., -1.5e-3, 2.), xtol=1e-14, ftol=1e-14, gtol=1e-14)
    np.testing.assert_allclose(woods(x, *par), woods(x, *popt), atol=1e-7)   # the curve; the parameters are strongly correlated
    np.testing.assert_allclose(par, popt, rtol=2e-3)
    assert chi == pytest.approx(((y - woods(x, *par)) ** 2).sum() / (x.size - 4), rel=1e-9)
    assert chi < 1.5                                                         # the reference's acceptance test (:863)
    sig = lambda x, a, b, c, d: a + (b - a) / (1 + (x / c) ** d)
    y = sig(x, 0.6, 2.9, 90.0, 0.45) + rng.normal(0, 0.004, x.size)
    par, chi = capi.graph_fit(3, x, y, (0.5, 3., 1e2, 0.4))                  # SetParameters(0.5, 3., 1e2, 0.4), :872
    popt, _ = curve_fit(sig, x, y, p0=(0.5, 3., 1e2, 0.4), xtol=1e-14, ftol=1e-14, gtol=1e-14, maxfev=20000)
    np.testing.assert_allclose(sig(x, *par), sig(x, *popt), atol=1e-6)
    assert chi == pytest.approx(((y - sig(x, *par)) ** 2).sum() / (x.size - 4), rel=1e-9)
    with pytest.raises(capi.NestB200Error) as e:
        capi.graph_fit(2, x[:4], y[:4], (10., 15., -1.5e-3, 2.))             # four points, four parameters
    assert e.value.code == capi.EFIT
    with pytest.raises(capi.NestB200Error):
        capi.graph_fit(0, x, y, (1., 1., 1., 1.))


def test_fit_covariance_and_propagated_leakage_error(built):
    """nb200_hist_fit_cov against scipy's covariance, and nb200_band_leakage_skew_cov (rootNEST.cpp:774-845) against the
    linear error propagation done numerically on the skew-normal CDF"""
    from scipy.optimize import curve_fit
    from scipy.special import erf
    from scipy.stats import skewnorm
    f = lambda x, A, xi, w, al: A / (w * np.sqrt(2 * np.pi)) * np.exp(-0.5 * ((x - xi) / w) ** 2) * (1 + erf(al * (x - xi) / (w * np.sqrt(2))))
    y = skewnorm.rvs(1.8, loc=2.45, scale=0.16, size=300000, random_state=3)
    mu, sd = y.mean(), y.std()
    lo, hi, bins = mu - 3 * sd, mu + 3 * sd, 60
    h, edges = np.histogram(y, bins=bins, range=(lo, hi))
    x = 0.5 * (edges[1:] + edges[:-1])
    start = (y.size * (hi - lo) / bins, mu - 0.5 * sd, 1.2 * sd, 0.5)
    par, cov, chi = capi.hist_fit_cov(1, h, lo, hi, start)
    par1, err1, chi1 = capi.hist_fit(1, h, lo, hi, start)
    assert np.array_equal(par, par1) and chi == chi1
    np.testing.assert_allclose(np.sqrt(np.diag(cov)), err1, rtol=1e-12)
    np.testing.assert_allclose(cov, cov.T, rtol=1e-9, atol=1e-18)
    k = h > 0
    popt, pcov = curve_fit(f, x[k], h[k].astype(float), p0=start, sigma=np.sqrt(h[k]), absolute_sigma=True, xtol=1e-14,
                           ftol=1e-14, gtol=1e-14)
    np.testing.assert_allclose(cov, pcov, rtol=5e-3, atol=1e-12)
    # leakage below a line 2.2 omega under xi, and its error
    line, line_err = par[1] - 2.2 * par[2] + 0.3, 0.004
    skew = np.array([[par[1], np.sqrt(cov[1, 1]), par[2], np.sqrt(cov[2, 2]), par[3], np.sqrt(cov[3, 3]), cov[1, 2], cov[1, 3],
                      cov[2, 3]]])
    leak, err = capi.band_leakage_skew_cov(skew, [line], [line_err])
    cdf = lambda xi, om, al, c: skewnorm.cdf(c, al, loc=xi, scale=om)
    assert leak[0] == pytest.approx(cdf(par[1], par[2], par[3], line), abs=2e-7)
    eps = 1e-6
    g = np.array([(cdf(par[1] + eps, par[2], par[3], line) - cdf(par[1] - eps, par[2], par[3], line)) / (2 * eps),
                  (cdf(par[1], par[2] + eps, par[3], line) - cdf(par[1], par[2] - eps, par[3], line)) / (2 * eps),
                  (cdf(par[1], par[2], par[3] + eps, line) - cdf(par[1], par[2], par[3] - eps, line)) / (2 * eps)])
    gc = (cdf(par[1], par[2], par[3], line + eps) - cdf(par[1], par[2], par[3], line - eps)) / (2 * eps)
    var = g @ cov[1:, 1:] @ g + (gc * line_err) ** 2
    assert err[0] == pytest.approx(np.sqrt(var), rel=2e-3)
    with pytest.raises(ValueError):
        capi.band_leakage_skew_cov(skew[:, :8], [line], [line_err])
