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
