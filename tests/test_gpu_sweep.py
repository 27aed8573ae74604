"""nb200_run_sweep (BASELINE config 4: 50 GeV WIMP + 8B over a grid of drift fields x g1 x g1_gas, one band per grid
point, all points in the same launches): bit-equal to per-point nb200_run, and against the compiled reference at the
same (field, g1, g1_gas)."""
import ctypes as C
import itertools

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import band_chi2, ref_chain
from nest_b200 import capi, shard

pytestmark = pytest.mark.gpu


def grid_points(fields, g1s, g1gs, kind, wimp=None):
    """one (detector, model, source, run constants) per grid point"""
    dets, mods, srcs, rcs, keys = [], [], [], [], []
    for f, g1, g1g in itertools.product(fields, g1s, g1gs):
        det = capi.detector("LUX_Run03")
        det.g1, det.g1_gas = g1, g1g
        rc = capi.prepare(det, f)         # host: density, drift speed, CalculateG2 for this point
        if kind == "wimp":
            src = wimp.attach(capi.source(capi.SRC_WIMP, capi.WIMP, 50.0, 1e-36, f))
        elif kind == "b8":
            src = capi.source(capi.SRC_B8, capi.B8, 0.0, 5.0, f)
        else:
            src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 60.0, f)
        dets.append(det)
        mods.append(capi.model(1))
        srcs.append(src)
        rcs.append(rc)
        keys.append((f, g1, g1g))
    return dets, mods, srcs, rcs, keys


@pytest.mark.parametrize("kind,n", [("wimp", 30_000), ("b8", 50_001), ("nr", 1024)])
def test_sweep_is_bit_equal_to_per_point_runs(ctx, kind, n):
    ana = capi.analysis(0)
    w = capi.WimpTable(50.0, 5.0, 0.0, seed=7)
    dets, mods, srcs, rcs, keys = grid_points((80.0, 180.0, 700.0), (0.09, 0.117), (0.08, 0.1), kind, w)
    bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=11, n=n, first_event=5)
    assert len(bands) == 12
    distinct = set()
    for k, b in enumerate(bands):
        one = capi.Band(ana)
        ctx.run(dets[k], mods[k], ana, srcs[k], rcs[k], 11, n, first_event=5, columns=[], band=one)
        assert np.array_equal(b.words(), one.words()), keys[k]
        assert b.c.n_events == n
        distinct.add(b.words().tobytes())
    assert len(distinct) == 12   # every grid point is its own physics


def test_sweep_spanning_several_workspace_chunks(ctx_small_chunks):
    """points x padded events beyond one workspace chunk (4.24e6 slots in this context): chunk boundaries fall inside
    points"""
    ctx = ctx_small_chunks
    ana = capi.analysis(0)
    dets, mods, srcs, rcs, keys = grid_points((180.0, 400.0), (0.117,), (0.1, 0.12, 0.14), "nr")
    n = 1_200_003
    bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=3, n=n)
    for k in (0, 3, 5):
        one = capi.Band(ana)
        ctx.run(dets[k], mods[k], ana, srcs[k], rcs[k], 3, n, columns=[], band=one)
        assert np.array_equal(bands[k].words(), one.words()), keys[k]
    # accumulating: a second sweep of the following event range adds to the same accumulators
    D, M, S, R = (capi.Detector * 6)(*dets), (capi.Model * 6)(*mods), (capi.Source * 6)(*srcs), (capi.RunConsts * 6)(*rcs)
    B = (capi.BandHist * 6)(*[b.c for b in bands])
    ctx._check(capi.lib().nb200_run_sweep(ctx.h, 6, D, M, C.byref(ana), S, R, 3, n, 1000, B))
    one = capi.Band(ana)
    ctx.run(dets[5], mods[5], ana, srcs[5], rcs[5], 3, n + 1000, columns=[], band=one)
    assert np.array_equal(bands[5].words(), one.words())


def test_sweep_rejects_what_it_does_not_do(ctx):
    ana = capi.analysis(0)
    dets, mods, srcs, rcs, _ = grid_points((180.0,), (0.117,), (0.1,), "nr")
    srcs[0].vec_mode = 1
    with pytest.raises(capi.NestB200Error) as ei:
        ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=1, n=100)
    assert ei.value.code == capi.EINVAL


POINTS = [(180.0, 0.117, 0.1), (60.0, 0.09, 0.12), (500.0, 0.14, 0.08), (1000.0, 0.117, 0.13)]


def test_sweep_points_match_reference(ctx, ref):
    """four grid points of the (field, g1, g1_gas) sweep, WIMP 50 GeV and 8B, against the reference run with the same
    detector settings (VDetector setters, as loopNEST): event-level signals by KS, the band by the reference's own
    GetBand + rootNEST's chi^2, the fraction of events inside the S1/S2 window"""
    ana = capi.analysis(0)
    n = 60_000
    seed_ref = 21
    iso = ref.sampler(6, seed_ref, 1_000_000 + 1100).astype(np.int32)
    w = capi.WimpTable(50.0, 5.0, 0.0, isotopes=iso)   # the table the reference builds under seed_ref
    try:
        for kind, sp, sk, e0, e1 in (("wimp", capi.WIMP, capi.SRC_WIMP, 50.0, 1e-36), ("b8", capi.B8, capi.SRC_B8, 0.0, 5.0)):
            dets, mods, srcs, rcs, keys = [], [], [], [], []
            for f, g1, g1g in POINTS:
                d, m, s, r, k = grid_points((f,), (g1,), (g1g,), kind, w)
                dets += d; mods += m; srcs += s; rcs += r; keys += k
            bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=5, n=n)
            for k, (f, g1, g1g) in enumerate(POINTS):
                ref.lux_set(g1, g1g)
                r = ref_chain(ref, seed_ref, n, sp, sk, e0, e1, inField=f)
                s1r, s2r = r[:, refprobe.SIG1], r[:, refprobe.SIG2]
                cols = ctx.run(dets[k], mods[k], ana, srcs[k], rcs[k], 5, n, columns=["signal1", "signal2", "s2_8"])
                # g2 of this grid point: the host's CalculateG2 against the reference's (ionization[8])
                assert abs(cols["s2_8"][0] - r[0, refprobe.S2_0 + 8]) <= 1e-12 * abs(r[0, refprobe.S2_0 + 8])
                for a, b in ((cols["signal1"], s1r), (cols["signal2"], s2r)):
                    assert stats.ks_2samp(a, b).pvalue > 1e-4, (kind, POINTS[k])
                bg = bands[k]
                # events inside the band window, counted on the reference's signals by the same GetBand rules
                wr = shard.band_words_from_signals(ana, dets[k].coinLevel, s1r, s2r)
                acc_g, acc_r = int(bg.accepted.sum()), int(wr[:ana.numBins * 6].reshape(-1, 6)[:, 0].sum())
                assert abs(acc_g - acc_r) < 6 * np.sqrt(max(acc_g + acc_r, 50)), (kind, POINTS[k], acc_g, acc_r)
                bref = ref.getband(s1r, s2r)
                if kind == "wimp":
                    b = bg.finalize()
                    nr_bin = wr[:ana.numBins * 6].reshape(-1, 6)[:, 0]
                    ok = (bg.accepted > 150) & (nr_bin > 150) & np.isfinite(bref).all(axis=1) & (bref[:, 4] > 0)
                    assert ok.sum() >= 8
                    cm, cw, dof = band_chi2(b[ok], bref[ok])
                    assert cm < 2.5 and cw < 2.5, (POINTS[k], cm, cw, dof)
    finally:
        ref.lux_set()


def test_sweep_band_post_on_device_equals_the_host_functions(ctx):
    """statistics, rootNEST's goodness of fit and leakage of a sweep's bands computed on the device (k_band_post)
    against nb200_band_finalize / nb200_band_chi2 / nb200_band_leakage on the copied-back accumulators"""
    ana = capi.analysis(0)
    dets, mods, srcs, rcs, keys = grid_points((100.0, 180.0, 400.0), (0.10, 0.117), (0.09, 0.1), "nr")
    n = 400_000
    bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=2, n=n)
    target = bands[3].finalize()                      # one grid point's band plays the data band
    ok = np.isfinite(target).all(axis=1)
    centroid = np.where(ok, target[:, 2] - 0.05, 1.0)
    stats_d, gof_d, leak_d = ctx.sweep_band_post(len(bands), ana, target=target, free_param=2, centroid=centroid)
    for k, b in enumerate(bands):
        st = b.finalize()
        assert np.array_equal(np.isnan(st), np.isnan(stats_d[k]))
        m = np.isfinite(st)
        assert np.max(np.abs(st[m] - stats_d[k][m]) / np.maximum(np.abs(st[m]), 1e-300)) < 1e-12
        chi = capi.band_chi2(st, target, useS2=0, free_param=2)
        assert np.allclose(gof_d[k], chi, rtol=1e-10, atol=1e-12), (keys[k], gof_d[k], chi)
        _, _, _, tot = b.leakage(centroid)
        assert np.allclose(leak_d[k], tot, rtol=1e-12), (keys[k], leak_d[k], tot)
    assert gof_d[3][0] == 0.0 and gof_d[3][1] == 0.0      # the target against itself
    assert np.argmin(gof_d[:, 2]) == 3


def _skew_window_fit_on_host(ana, band, j):
    """the skew-Gaussian fit k_band_fit makes for slice j, redone with the host entry points: GetBand's moments
    (nb200_band_finalize), the window mean +- 3 sigma on the fixed log axis, EstimateSkew's start values
    (examples/rootNEST.cpp:1567-1585, 1367-1374), nb200_hist_fit_cov(1, ...)"""
    st = band.finalize()
    lb, N = ana.logBins, float(band.accepted[j])
    w = (ana.logMax - ana.logMin) / lb
    mean, sd = st[j, 2], st[j, 3]
    if not (sd > 0 and np.isfinite(mean) and np.isfinite(st[j, 6])):
        return None
    skew = st[j, 6] * (N - 2.0) / N
    lo, hi = mean - 3 * sd, mean + 3 * sd
    if lo > 2.0 or hi > 3.0:
        lo, hi = lo - 0.5, hi + 0.5
    x = ana.logMin + (np.arange(lb) + 0.5) * w
    inw = np.nonzero((x >= lo) & (x < hi))[0]
    if inw.size == 0:
        return None
    k0, k1 = int(inw[0]), int(inw[-1]) + 1
    row = band.hist2d.reshape(ana.numBins, lb)[j, k0:k1]
    n = float(row.sum())
    if not n > 1:
        return None
    if skew == 0.0:
        skew = 1e-9
    skew = np.sign(skew) * min(0.7, abs(skew))
    s23 = abs(skew) ** (2.0 / 3.0)
    delta0 = np.sign(skew) * np.sqrt(np.pi / 2 * s23 / (s23 + ((4 - np.pi) / 2) ** (2.0 / 3.0)))
    alpha0 = delta0 / np.sqrt(1 - delta0 * delta0)
    dE = alpha0 / np.sqrt(1 + alpha0 * alpha0)
    omega0 = sd / np.sqrt(1 - 2 * dE * dE / np.pi)
    start = [n * w, mean - omega0 * dE * np.sqrt(2 / np.pi), omega0, alpha0]
    try:
        par, cov, chi = capi.hist_fit_cov(1, row, ana.logMin + k0 * w, ana.logMin + k1 * w, start)
    except capi.NestB200Error:
        return None
    return par, cov, chi, int(np.count_nonzero(row))


def test_per_bin_band_fits_on_device_equal_the_host_functions(ctx):
    """the Gaussian and skew-Gaussian fits of every S1 slice of every grid point's band and the leakage
    read off them, computed on the device (k_band_fit: nb_fit.h's statements, one thread per slice), against the host
    entry points on the copied-back accumulators (nb200_band_fit_gauss, nb200_hist_fit_cov, nb200_band_leakage_gauss /
    _skew)"""
    ana = capi.analysis(0)
    dets, mods, srcs, rcs, keys = grid_points((100.0, 180.0, 400.0), (0.10, 0.117), (0.1,), "nr")
    bands = ctx.run_sweep(dets, mods, ana, srcs, rcs, seed=4, n=400_000)
    nr_line = bands[1].finalize()[:, 2]
    line = np.where(np.isfinite(nr_line), nr_line - 0.1, 1.0)
    # Gaussian
    fit_d, leak_d = ctx.sweep_band_fit(len(bands), ana, model=0, line=line)
    fitted = 0
    for k, b in enumerate(bands):
        fh = b.fit_gauss()
        assert np.array_equal(fh[:, 7], fit_d[k][:, 7]), keys[k]               # same slices fitted, same points used
        m = fh[:, 7] > 0
        fitted += int(m.sum())
        assert np.allclose(fit_d[k][m][:, :5], fh[m][:, :5], rtol=1e-7, atol=1e-10), keys[k]
        assert np.allclose(fit_d[k][m][:, 5:7], fh[m][:, 5:7], rtol=1e-6, atol=1e-9), keys[k]
        nr = np.zeros_like(fh)
        nr[:, 1] = line
        lk = capi.band_leakage_gauss(fh, nr, 0.0)[:, 1]
        assert np.allclose(leak_d[k][m], lk[m], rtol=1e-6, atol=1e-12), keys[k]
        assert np.all(leak_d[k][~m] == 0.0)
    assert fitted > 300
    # skew-Gaussian
    fit_s, leak_s = ctx.sweep_band_fit(len(bands), ana, model=1, line=line)
    checked = 0
    for k in (0, 3, 5):
        b = bands[k]
        for j in range(0, ana.numBins, 5):
            h = _skew_window_fit_on_host(ana, b, j)
            if h is None:
                assert fit_s[k][j, 11] == 0.0, (keys[k], j)
                continue
            par, cov, chi, used = h
            o = fit_s[k][j]
            assert o[11] == used, (keys[k], j, o[11], used)
            assert np.allclose(o[:4], par, rtol=1e-6, atol=1e-9), (keys[k], j, o[:4], par)
            assert np.allclose(o[4:7], np.sqrt(np.diag(cov))[1:], rtol=1e-5), (keys[k], j)
            assert np.allclose(o[7:10], [cov[1, 2], cov[1, 3], cov[2, 3]], rtol=1e-4, atol=1e-12), (keys[k], j)
            assert abs(o[10] - chi) <= 1e-6 * max(chi, 1.0), (keys[k], j, o[10], chi)
            lk, _, _ = capi.band_leakage_skew([par[1]], [par[2]], [0.0], [par[3]], [line[j]])
            assert abs(leak_s[k][j] - lk[0]) < 1e-7, (keys[k], j, leak_s[k][j], lk[0])
            checked += 1
    assert checked >= 30
    # the context's own accumulator (nb200_run_async) goes through the same kernel
    ctx.band_reset(ana)
    ctx.run_async(dets[1], mods[1], ana, srcs[1], rcs[1], 4, 0, 400_000)
    one, _ = ctx.band_fit_device(ana, model=0)
    assert np.array_equal(one, fit_d[1])
    with pytest.raises(capi.NestB200Error):
        ctx.sweep_band_fit(len(bands) + 1, ana)            # no sweep of that shape
    with pytest.raises(capi.NestB200Error):
        ctx.sweep_band_fit(len(bands), ana, model=2)
