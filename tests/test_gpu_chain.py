"""Statistical parity of the stochastic chain on the GPU against the compiled reference.

The GPU path replaces the reference's sequential xoroshiro stream by Philox(seed, event), so
equal seeds do not give equal numbers: every output column is compared as a distribution
(two-sample Kolmogorov-Smirnov) and the band by the reference's own chi^2
(examples/rootNEST.cpp:619-643).  Fixed seeds make the outcome reproducible."""
import numpy as np
import pytest

import refprobe
from helpers import COLUMN_NAMES, band_chi2, gpu_chain, ks_report, mean_z, ref_chain
from nest_b200 import capi

pytestmark = pytest.mark.gpu
P_MIN = 1e-4   # per-column KS p-value floor (about 40 columns x 10 configurations)


def check(a, b, skip=(), label=""):
    ks = ks_report(a, b, skip)
    bad = {k: v for k, v in ks.items() if v < P_MIN}
    z = mean_z(a, b)
    assert not bad, "%s: KS failures %r (mean z-scores %r)" % (label, bad, {k: z[k] for k in bad})


MONO = [
    ("beta_5keV", capi.beta, 5.0, 0, 1.0, 1),
    ("beta_0.7keV", capi.beta, 0.7, 0, 1.0, 1),
    ("beta_50keV_defaultwidths_skew", capi.beta, 50.0, -1, 1.0, 0),
    ("gamma_41keV", capi.gammaRay, 41.5, -1, 1.0, 1),
    ("NR_3keV", capi.NR, 3.0, -1, 1.0, 1),
    ("NR_30keV", capi.NR, 30.0, -1, 1.0, 1),
    ("NR_250keV_parametric", capi.NR, 250.0, -1, 1.0, 1),
    ("alpha_5300keV", capi.ion, 5300.0, -1, 1.0, 1),
]


@pytest.mark.parametrize("label,species,E,er_model,mult,which", MONO)
def test_mono_fixed_position(ctx, ref, label, species, E, er_model, mult, which):
    """mono-energetic source at a fixed vertex: isolates quanta, S1 and S2 distributions"""
    n = 60000
    xyz = (70.0, -40.0, 300.0)
    g, _ = gpu_chain(ctx, 11, n, species, capi.SRC_MONO, E, E, 180.0, 1, xyz, er_model, mult, which)
    r = ref_chain(ref, 7, n, species, capi.SRC_MONO, E, E, 180.0, 1, xyz, er_model, mult, which)
    check(g, r, label=label)


@pytest.mark.parametrize("label,species,kind,emin,emax,er_model", [
    ("config1_ER_flat", capi.beta, capi.SRC_FLAT, 0.1, 20.0, 0),
    ("config2_NR_flat", capi.NR, capi.SRC_FLAT, 0.0, 100.0, -1),
    ("CH3T", capi.CH3T, capi.SRC_CH3T, 0.0, 18.5898, -1),
    ("DD", capi.DD, capi.SRC_DD, 0.0, 80.0, -1),
    ("B8", capi.B8, capi.SRC_B8, 0.0, 5.0, -1),
    ("AmBe", capi.AmBe, capi.SRC_AMBE, 0.5, 300.0, -1),
    ("Cf", capi.Cf, capi.SRC_CF, 0.0, 200.0, -1),
    ("C14", capi.C14, capi.SRC_C14, 0.0, 156.0, -1),
    ("ppSolar", capi.ppSolar, capi.SRC_PPSOLAR, 0.0, 250.0, -1),
    ("atmNu", capi.atmNu, capi.SRC_ATMNU, 0.0, 85.0, -1),
])
def test_random_vertex_chain_and_band(ctx, ref, label, species, kind, emin, emax, er_model):
    n = 100000
    par = (13., 0.12, 71.2, 20., -20.5) if kind == capi.SRC_DD else (-0.95351,) if kind == capi.SRC_AMBE else ()
    ana = capi.analysis(0)
    band = capi.Band(ana)
    g, cols = gpu_chain(ctx, 3, n, species, kind, emin, emax, 180.0, 0, (0, 0, 0), er_model, 1.0, 1, ana=ana,
                        band=band, par=par)
    r = ref_chain(ref, 5, n, species, kind, emin, emax, 180.0, 0, (0, 0, 0), er_model, 1.0, 1)
    check(g, r, label=label)
    # band: GPU accumulator vs the reference's GetBand on the reference's own signals
    bg = band.finalize()
    br = ref.getband(r[:, refprobe.SIG1], r[:, refprobe.SIG2])
    if label != "B8":
        cm, cw, dof = band_chi2(bg, br)
        assert cm < 2.0 and cw < 2.0, "%s band chi2/dof mean %.2f width %.2f (dof %d)" % (label, cm, cw, dof)
    # and the GPU accumulator reproduces GetBand on the GPU's own signals (same events)
    bs = ref.getband(cols["signal1"], cols["signal2"])
    ok = np.isfinite(bs).all(axis=1) & np.isfinite(bg).all(axis=1)
    assert np.allclose(bg[ok][:, [0, 1, 2, 3, 4, 5]], bs[ok][:, [0, 1, 2, 3, 4, 5]], rtol=2e-6, atol=1e-7)
    # column 6, the skewness (execNEST.cpp:1612-1618): the third central moment from the 2^-22 fixed-point sum of y^3
    # against the reference's sum of cubed deviations -- bins with enough events for the cancellation to be benign
    pop = ok & (band.accepted > 300)
    # sources that put most of their events above the band's S1 range (or below threshold) leave no such bin
    assert pop.sum() >= 1 or label in ("B8", "C14", "ppSolar", "atmNu", "AmBe", "Cf")
    assert np.allclose(bg[pop][:, 6], bs[pop][:, 6], rtol=2e-3, atol=2e-3), np.abs(bg[pop][:, 6] - bs[pop][:, 6]).max()


@pytest.mark.parametrize("line,maxT", [(41.5, 1000.0), (9.4, 400.0), (32.1, 400.0)])
def test_kr83m_chain(ctx, ref, line, maxT):
    """83mKr: eMin = the line, eMax = max time between the two decays (execNEST.cpp:583-588);
    the separation is drawn per event and shapes the 9.4 keV yields (NEST.cpp:954-1044)"""
    n = 60000
    g, cols = gpu_chain(ctx, 31, n, capi.Kr83m, capi.SRC_MONO, line, maxT)
    r = ref_chain(ref, 32, n, capi.Kr83m, capi.SRC_MONO, line, maxT)
    check(g, r, label="Kr83m %g keV" % line)
    dT = cols["deltaT_ns"]
    assert dT.min() >= 0.0 and dT.max() <= maxT
    from scipy import stats
    exp = ref.sampler(5, 77, n, 154.4, 0.0, maxT)   # rand_exponential(half-life, min, max)
    assert stats.ks_2samp(dT, exp).pvalue > P_MIN


def test_nonuniform_field(ctx, ref):
    n = 40000
    g, _ = gpu_chain(ctx, 21, n, capi.beta, capi.SRC_FLAT, 1.0, 15.0, -1.0, 0, (0, 0, 0), 0, 1.0, 1)
    r = ref_chain(ref, 22, n, capi.beta, capi.SRC_FLAT, 1.0, 15.0, -1.0, 0, (0, 0, 0), 0, 1.0, 1)
    check(g, r, label="FitEF field")


def test_event_results_do_not_depend_on_batching(ctx):
    """event e depends only on (seed, e): shards and chunk sizes reproduce the same records"""
    n = 5000
    a, _ = gpu_chain(ctx, 9, n, capi.NR, capi.SRC_FLAT, 0.0, 100.0)
    b1, _ = gpu_chain(ctx, 9, 1777, capi.NR, capi.SRC_FLAT, 0.0, 100.0, first_event=0)
    b2, _ = gpu_chain(ctx, 9, n - 1777, capi.NR, capi.SRC_FLAT, 0.0, 100.0, first_event=1777)
    assert np.array_equal(a, np.vstack([b1, b2]))
    c, _ = gpu_chain(ctx, 10, n, capi.NR, capi.SRC_FLAT, 0.0, 100.0)
    assert not np.array_equal(a, c)


def test_band_accumulator_is_exact(ctx):
    """integer band words == recomputation from the per-event signals; shards add up exactly"""
    ana = capi.analysis(0)
    n = 30000
    full = capi.Band(ana)
    _, cols = gpu_chain(ctx, 4, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0, ana=ana, band=full)
    parts = capi.Band(ana)
    for lo, hi in [(0, 7000), (7000, 7001), (7001, 30000)]:
        gpu_chain(ctx, 4, hi - lo, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0, ana=ana, band=parts,
                  first_event=lo)
    assert np.array_equal(full.words(), parts.words())
    s1, s2 = cols["signal1"], cols["signal2"]
    w = (ana.maxS1 - ana.minS1) / ana.numBins
    acc = np.zeros(ana.numBins, np.int64)
    rej = np.zeros(ana.numBins, np.int64)
    for j in range(ana.numBins):
        c = ana.minS1 + w / 2. + j * w
        m = (np.abs(s1) > c - w / 2.) & (np.abs(s1) <= c + w / 2.)
        good = m & (s1 >= 0) & (s2 >= 0)
        acc[j] = good.sum()
        rej[j] = (m & ~good).sum()
    assert np.array_equal(acc, full.accepted.astype(np.int64))
    assert np.array_equal(rej, full.rejected.astype(np.int64))
    assert full.hist2d.sum() <= full.accepted.sum()


def test_edge_cases(ctx):
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo, ana = capi.model(1), capi.analysis(0)
    # no events: nothing to do, no error
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0)
    assert ctx.run(det, mo, ana, src, rc, 1, 0) is not None
    # energy below the work function: zero quanta, zero signals (execNEST.cpp:1066-1077)
    g, cols = gpu_chain(ctx, 1, 100, capi.NR, capi.SRC_MONO, 0.005, 0.005)
    assert not cols["photons"].any() and not cols["electrons"].any()
    assert np.all(cols["signal1"] == -999.) and np.all(cols["signalE"] == 0.)
    # zero drift field: no S2 (NEST.cpp:1746-1750)
    g, cols = gpu_chain(ctx, 1, 100, capi.beta, capi.SRC_MONO, 10.0, 10.0, inField=0.5, fixed_pos=1,
                        xyz=(0., 0., 300.))
    assert not cols["s2_4"].any()
    # bad parameters are refused like the reference (return 1 + message)
    bad = capi.model(1)
    bad.NRERWidthsParam[9] = 5.0
    with pytest.raises(capi.NestB200Error) as e:
        ctx.run(det, bad, ana, capi.source(capi.SRC_FLAT, capi.beta, 0.1, 20.0), rc, 1, 10)
    assert e.value.code == capi.EPHYSICS
    ana2 = capi.analysis(0)
    ana2.numBins = 5000
    with pytest.raises(capi.NestB200Error):
        ctx.run(det, mo, ana2, src, rc, 1, 10)


def test_few_hit_s1_area_shape(ctx, ref):
    """the hit loop draws a photo-electron through its sum S = T + N with an exact accept test
    instead of the reference's two normals: compare the S1 area of 1-, 2- and 3-hit events at high
    statistics, where the single-phe shape (truncation, noise, double-phe, threshold) is exposed"""
    n = 400000
    xyz = (20.0, 30.0, 280.0)
    g, _ = gpu_chain(ctx, 21, n, capi.beta, capi.SRC_MONO, 1.5, 1.5, 180.0, 1, xyz, 0, 1.0, 1)
    r = ref_chain(ref, 22, n, capi.beta, capi.SRC_MONO, 1.5, 1.5, 180.0, 1, xyz, 0, 1.0, 1)
    S1 = refprobe.S1_0
    from scipy import stats
    for hits in (1, 2, 3):
        a = g[np.abs(g[:, S1 + 0]) == hits]
        b = r[np.abs(r[:, S1 + 0]) == hits]
        assert len(a) > 3000 and len(b) > 3000, (hits, len(a), len(b))
        assert abs(len(a) - len(b)) < 6 * np.sqrt(len(a))
        for col in (1, 2, 6):      # Nphe, pulse area, spike
            p = stats.ks_2samp(np.abs(a[:, S1 + col]), np.abs(b[:, S1 + col])).pvalue
            assert p > P_MIN, (hits, col, p)
        # fraction of hits below the sPE threshold / lost to the coincidence window
        fa, fb = (a[:, S1 + 2] == 0).mean(), (b[:, S1 + 2] == 0).mean()
        assert abs(fa - fb) < 5 * np.sqrt(fb * (1 - fb) / len(b) * 2 + 1e-12), (hits, fa, fb)


def test_double_phe_count_per_event_keeps_the_joint_law(ctx, ref):
    """the hit loop draws the NUMBER of double-phe hits once per event (the hits are exchangeable) instead of
    the reference's Bernoulli per hit (NEST.cpp:1391): the joint law of (nHits, Nphe, area, spike) must not
    move.  Conditional on nHits, compare Nphe - nHits ~ Binomial(nHits, P_dphe) exactly, and against the
    reference the covariance of the extra photo-electrons with the pulse area and the spike count."""
    from scipy import stats
    n = 300000
    xyz = (20.0, 30.0, 280.0)
    g, _ = gpu_chain(ctx, 31, n, capi.beta, capi.SRC_MONO, 6.0, 6.0, 180.0, 1, xyz, 0, 1.0, 1)
    r = ref_chain(ref, 32, n, capi.beta, capi.SRC_MONO, 6.0, 6.0, 180.0, 1, xyz, 0, 1.0, 1)
    S1 = refprobe.S1_0
    p_dphe = capi.detector("LUX_Run03").P_dphe

    def cols(a):
        a = a[np.abs(a[:, S1 + 0]) > 3]   # above the coincidence level: every hit is counted
        return np.abs(a[:, S1 + 0]), np.abs(a[:, S1 + 1]), np.abs(a[:, S1 + 2]), np.abs(a[:, S1 + 6])

    hg, pg, ag, sg = cols(g)
    hr, pr, ar, sr = cols(r)
    # exact conditional law of the extra photo-electrons, pooled over nHits through their z-scores
    for h, pe in ((hg, pg), (hr, pr)):
        extra = pe - h
        assert abs(extra.sum() - p_dphe * h.sum()) < 5 * np.sqrt(p_dphe * (1 - p_dphe) * h.sum())
        z = (extra - p_dphe * h) / np.sqrt(p_dphe * (1 - p_dphe) * h)
        assert abs(z.var() - 1) < 5 * np.sqrt(2 / z.size) + 0.01
    for hits in (20, 30, 40):
        a, b = (pg - hg)[hg == hits], (pr - hr)[hr == hits]
        assert len(a) > 2000 and len(b) > 2000
        obs = np.bincount(a.astype(int), minlength=hits + 1).astype(float)
        exp = stats.binom.pmf(np.arange(hits + 1), hits, p_dphe) * len(a)
        keep = exp >= 10
        o = np.concatenate([obs[keep], [obs[~keep].sum()]])
        e = np.concatenate([exp[keep], [exp[~keep].sum()]])
        assert stats.chisquare(o[e > 0], e[e > 0] * o.sum() / e[e > 0].sum()).pvalue > P_MIN, hits
    # joint moments against the reference: residuals after removing the dependence on nHits
    def resid(h, y):
        k = np.polyfit(h, y, 1)
        return y - np.polyval(k, h)
    for name, yg, yr in (("area", ag, ar), ("spike", sg, sr)):
        cg = np.mean(resid(hg, pg - hg) * resid(hg, yg))
        cr = np.mean(resid(hr, pr - hr) * resid(hr, yr))
        vg = np.var(resid(hg, pg - hg) * resid(hg, yg)) / hg.size
        vr = np.var(resid(hr, pr - hr) * resid(hr, yr)) / hr.size
        assert abs(cg - cr) < 5 * np.sqrt(vg + vr), (name, cg, cr)
        # a second photo-electron adds ~1 phe of area to its hit: the covariance is not zero
        if name == "area":
            assert cg > 10 * np.sqrt(vg)


def _variant_detectors():
    a = capi.detector("LUX_Run03")        # efficiency draws, 3-fold coincidence, S1 linear noise
    a.sPEeff, a.coinLevel, a.numPMTs = 0.85, 3, 61
    a.noiseLinear[0] = 0.03
    b = capi.detector("LUX_Run03")        # no baseline noise (tau = 0), S2 bottom array, quadratic noise
    b.noiseBaseline[1] = 0.0
    b.sPEres, b.P_dphe, b.s2_thr = 0.58, 0.2, -150.0
    b.noiseQuadratic[0], b.noiseQuadratic[1] = 1e-4, 1e-6
    c = capi.detector("LUX_Run03")        # large baseline noise with an offset, 1-fold, noise electrons
    # (a noise-electron draw that drives Ne negative sends the reference's std::binomial_distribution
    # into undefined behaviour, so keep the added electrons positive)
    c.noiseBaseline[0], c.noiseBaseline[1], c.noiseBaseline[2], c.noiseBaseline[3] = 0.05, 0.3, 30.0, 2.0
    c.coinLevel, c.sPEthr, c.coinWind = 1, 0.5, 30.0
    return {"eff85_3fold": a, "nonoise_s2bottom": b, "bignoise_1fold": c}


@pytest.mark.parametrize("which", ["eff85_3fold", "nonoise_s2bottom", "bignoise_1fold"])
def test_other_detector_parameters_against_oracle(ctx, orc, which):
    """detector settings that exercise the rarely-taken branches (sPEeff < 1, nCr coincidence,
    S2-bottom, quadratic noise, zero / large baseline noise): the oracle port takes any flattened
    detector (it is pinned bit-for-bit to the reference on LUX_Run03)"""
    det = _variant_detectors()[which]
    capi.prepare(det, 180.0)
    n = 80000
    saved = orc.det
    try:
        orc.det = det
        for species, kind, e0, e1, erm in [(capi.beta, capi.SRC_FLAT, 0.1, 12.0, 0), (capi.NR, capi.SRC_FLAT, 0.0, 60.0, -1)]:
            g, _ = gpu_chain(ctx, 31, n, species, kind, e0, e1, er_model=erm, det=det)
            r = ref_chain(orc, 32, n, species, kind, e0, e1, er_model=erm)
            check(g, r, label="%s species %d" % (which, species))
    finally:
        orc.det = saved
