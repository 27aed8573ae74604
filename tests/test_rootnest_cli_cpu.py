"""rootNEST_b200 (nest_b200/host/rootNEST_main.cpp): rootNEST's band-of-Gaussians, leakage and goodness-of-fit
modes without ROOT, on execNEST's 12-column stdout.  No GPU: the tool only post-processes text.  Checked
against numpy restatements of reference examples/rootNEST.cpp (GetFile :955-1036, GetBand :1198-1306,
leakage :668-950, GoF :587-666) on the same rows, and run on the unmodified reference binary's own output."""
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.special import erf

import refprobe
from nest_b200 import capi, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "nest_b200", "host", "rootNEST_b200")


@pytest.fixture(scope="module")
def tool(built):
    r = subprocess.run(["make", "-s", "-C", os.path.dirname(TOOL)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return TOOL


def run(tool, *args, env=None):
    e = dict(os.environ)
    e["NB200_SKEWNESS"] = "0"                 # Gaussian fits and per-bin NR means unless a test asks for the reference's defaults
    e["NB200_NRBANDCENTER"] = "1"
    e.update(env or {})
    r = subprocess.run([tool] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=300)
    return r.returncode, r.stdout, r.stderr


def write_execnest_file(path, seed, n, mu0, slope, sigma, alpha=0.0, s1max=70.0):
    """rows in execNEST's stdout layout (src/execNEST.cpp:1386-1405) under its two header lines; some rows carry the
    reference's below-threshold flags (negative signals) and some lie outside the S1 / S2 windows"""
    from scipy.stats import skewnorm
    rng = np.random.default_rng(seed)
    s1 = rng.uniform(0.5, s1max, n)
    s2 = s1 * 10.0 ** (mu0 + slope * s1 + skewnorm.rvs(alpha, scale=sigma, size=n, random_state=rng))
    s1[::17] *= -1.0
    s2[::23] *= -1.0
    s1, s2 = np.round(s1, 6), np.round(s2, 6)       # what the 6-decimal text carries
    with open(path, "w") as f:
        f.write("E_truth [keV]\tfield [V/cm]\ttDrift [us]\tX,Y,Z [mm]\tNph\tNe-\tS1 [PE or phe]\tS1_3Dcor [phd]\t"
                "spikeC(NON-INT)\tNe-Extr\tS2_rawArea [PE]\tS2_3Dcorr [phd]\n")
        for a, b in zip(s1, s2):
            f.write("%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d\t%.6f\t%.6f\t%.6f\t%d\t%.6f\t%.6f\n" %
                    (5.0, 180.0, 100.0, 10, -20, 250, 300, 200, a * 1.1, a * 0.97, a, 100, b * 0.8, b))
    return s1, s2


def windowed(ana, s1, s2):
    """GetFile's flags (:996-1025, usePD = 2): outside (minS, maxS) -> -999"""
    a = np.where((np.abs(s1) > ana.minS1) & (s1 < ana.maxS1), s1, -999.0)
    b = np.where((np.abs(s2) > ana.minS2) & (s2 < ana.maxS2), s2, -999.0)
    return a, b


def table(out, after=None):
    lines = out.splitlines()
    if after is not None:
        lines = lines[lines.index(after):]
    k = next(i for i, ln in enumerate(lines) if ln.startswith("Bin Center\tBin Actual\tGaus Mean"))
    rows = []
    for ln in lines[k + 1:]:
        f = ln.split("\t")
        if len(f) != 9:
            break
        rows.append([float(v) for v in f])
    return np.array(rows)


def test_band_of_gaussians(tool, tmp_path):
    ana = capi.analysis(0)
    s1, s2 = write_execnest_file(tmp_path / "er.txt", 1, 120000, 2.6, -0.006, 0.11)
    rc, out, err = run(tool, str(tmp_path / "er.txt"))
    assert rc == 0, err
    t = table(out)
    assert t.shape == (ana.numBins, 9)
    w1, w2 = windowed(ana, s1, s2)
    band = shard.band_from_words(ana, shard.band_words_from_signals(ana, 2, w1, w2))
    fit, mom = band.fit_gauss(), band.finalize()
    np.testing.assert_allclose(t[:, 0], mom[:, 0], atol=1e-6)
    filled = band.accepted > 100
    assert filled.sum() > 60
    np.testing.assert_allclose(t[filled, 1], mom[filled, 1], atol=2e-6)        # Bin Actual = mean S1 of the bin
    for col, k in ((2, 1), (3, 3), (4, 2), (5, 4), (8, 6)):                   # mean, its error, sigma, its error, X^2/DOF
        np.testing.assert_allclose(t[filled, col], fit[filled, k], atol=2e-6)
    assert np.all(t[:, 6] == 0) and np.all(t[:, 7] == 0)                       # Gaussian fits: no skew columns (:1343-1345)
    # the fitted band is the one that was written
    low = filled & (t[:, 0] < 25)                       # above, the S2 window (maxS2 = 1e4) cuts into the band
    assert np.abs(t[low, 2] - (2.6 - 0.006 * t[low, 0])).max() < 0.02 and np.abs(t[low, 4] - 0.11).max() < 0.02


def test_leakage_table(tool, tmp_path):
    ana = capi.analysis(0)
    e1, e2 = write_execnest_file(tmp_path / "er.txt", 2, 150000, 2.6, -0.006, 0.11)
    n1, n2 = write_execnest_file(tmp_path / "nr.txt", 3, 80000, 2.3, -0.003, 0.12)
    rc, out, err = run(tool, str(tmp_path / "er.txt"), str(tmp_path / "nr.txt"))      # examples/leakage.sh's call
    assert rc == 0, err
    assert "Calculating band for first dataset" in out and "Calculating band for second dataset" in out
    nr_t = table(out, "Calculating band for first dataset")       # argv[2] is read first (:668-676)
    er_t = table(out, "Calculating band for second dataset")
    lines = err.splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("#evts\tBinCen"))
    assert "Leak Frac Raw" in lines[k]                             # ER given first: the ER-into-NR table
    rows = np.array([[float(v) for v in ln.split("\t")[:12]] for ln in lines[k + 1:k + 1 + ana.numBins]])
    assert rows.shape == (ana.numBins, 12)
    # counting, restated: events of the ER file per S1 bin below the NR Gaussian mean of that bin
    w1, w2 = windowed(ana, e1, e2)
    width = (ana.maxS1 - ana.minS1) / ana.numBins
    ok = (w1 >= 0) & (w2 >= 0)
    tot_below = tot = 0
    for i in range(ana.numBins):
        lo = ana.minS1 + i * width
        m = ok & (w1 > lo) & (w1 <= lo + width)
        y = np.log10(w2[m] / w1[m])
        assert rows[i, 0] == m.sum()
        if m.sum() == 0:
            continue
        below = (y < nr_t[i, 2]).sum()                             # the table prints 6 decimals of the centroid
        assert abs(rows[i, 8] * m.sum() - below) <= 1 + (np.abs(y - nr_t[i, 2]) < 1e-6).sum()
        N = float(m.sum())
        kk = rows[i, 8] * N
        assert rows[i, 9] == pytest.approx((kk + np.sqrt(kk)) / (N - np.sqrt(N)) - rows[i, 8], rel=1e-4, abs=1e-9)
        assert rows[i, 11] == pytest.approx((1 - rows[i, 8]) * 100, abs=1e-5)
        tot_below += round(kk)
        tot += m.sum()
        if er_t[i, 4] > 0 and nr_t[i, 4] > 0:
            n_sig = (er_t[i, 2] - nr_t[i, 2]) / er_t[i, 4]          # :719
            assert rows[i, 3] == pytest.approx(n_sig, abs=2e-4)
            assert rows[i, 4] == pytest.approx((1 - erf(rows[i, 3] / np.sqrt(2))) / 2, rel=1e-4)
            assert rows[i, 7] == pytest.approx((1 - rows[i, 4]) * 100, abs=1e-5)
    overall = [ln for ln in lines if ln.startswith("OVERALL DISCRIMINATION (ER)")][0]
    assert float(overall.split("Leakage Fraction = ")[1]) == pytest.approx(tot_below / tot, rel=1e-9)
    # NR first: the other header (characterising the NR band against its own centroid: about half below)
    rc, out, err = run(tool, str(tmp_path / "nr.txt"), str(tmp_path / "er.txt"))
    assert rc == 0 and "Lower ''Half''" in err
    overall = [ln for ln in err.splitlines() if ln.startswith("OVERALL DISCRIMINATION (ER)")][0]
    assert 0.45 < float(overall.split("Leakage Fraction = ")[1]) < 0.55


def test_leakage_below_a_curve_through_the_nr_means(tool, tmp_path):
    """NRbandCenter = 3 (the reference's default, analysis.hh:76) and 2: the Woods function fitted through the NR means
    (rootNEST.cpp:855-862), evaluated at each event's S1 (:883-886) or at the bin centre (:890-893)"""
    ana = capi.analysis(0)
    e1, e2 = write_execnest_file(tmp_path / "er.txt", 12, 150000, 2.6, -0.006, 0.11, s1max=40.0)
    write_execnest_file(tmp_path / "nr.txt", 13, 80000, 2.3, -0.003, 0.12, s1max=40.0)
    w1, w2 = windowed(ana, e1, e2)
    ok = (w1 >= 0) & (w2 >= 0) & (w1 > ana.minS1) & (w1 <= ana.maxS1)
    y = np.log10(w2[ok] / w1[ok])
    width = (ana.maxS1 - ana.minS1) / ana.numBins
    results = {}
    for center in ("1", "2", "3"):
        rc, out, err = run(tool, str(tmp_path / "er.txt"), str(tmp_path / "nr.txt"), env={"NB200_NRBANDCENTER": center})
        assert rc == 0 and "Poor fit" not in err, err
        nr_t = table(out, "Calculating band for first dataset")
        overall = [ln for ln in err.splitlines() if ln.startswith("OVERALL DISCRIMINATION (ER)")][0]
        results[center] = float(overall.split("Leakage Fraction = ")[1])
    good = nr_t[:, 4] > 0
    par, chi = capi.graph_fit(2, nr_t[good, 0], nr_t[good, 2], (10., 15., -1.5e-3, 2.))
    assert chi < 1.5
    woods = lambda x: par[0] / (x + par[1]) + par[2] * x + par[3]
    assert results["3"] == pytest.approx((y < woods(w1[ok])).mean(), rel=0.01, abs=2e-5)
    centre = ana.minS1 + (np.floor((w1[ok] - ana.minS1) / width - 1e-12) + 0.5) * width
    assert results["2"] == pytest.approx((y < woods(centre)).mean(), rel=0.02, abs=3e-5)
    # the three definitions of the line agree on the overall leakage to within tens of per cent, not digit for digit
    assert 0.5 < results["3"] / results["1"] < 1.5 and results["3"] != results["1"]


def test_goodness_of_fit_mode(tool, tmp_path):
    ana = capi.analysis(0)
    # S1 below 25 keeps the band clear of the S2 window; 120 log bins so that a bin is a fraction of the band's sigma
    fine = {"NB200_LOGBINS": "120"}
    write_execnest_file(tmp_path / "a.txt", 4, 150000, 2.6, -0.006, 0.11, s1max=25.0)
    write_execnest_file(tmp_path / "b.txt", 5, 150000, 2.6, -0.006, 0.11, s1max=25.0)
    rc, out, _ = run(tool, str(tmp_path / "b.txt"), env=fine)
    t = table(out)
    # a "data band" file in the layout rootNEST reads (:603-607): centre, actual, mean, mean error, sigma, sigma error
    sel = t[:, 4] > 0
    with open(tmp_path / "band.txt", "w") as f:
        for r in t:
            if r[4] > 0:
                f.write("%f\t%f\t%f\t%f\t%f\t%f\n" % (r[0], r[1], r[2], r[3], r[4], r[5]))
            else:
                f.write("%f\t%f\t%f\t%f\t%f\t%f\n" % (r[0], r[0], 2.0, 1e9, 0.1, 1e9))   # an empty bin: no weight
    rc, out, err = run(tool, str(tmp_path / "a.txt"), str(tmp_path / "band.txt"), env={"NB200_MODE": "1", **fine})
    assert rc == 0, err
    line = [ln for ln in out.splitlines() if ln.startswith("The reduced CHI^2 = ")][0]
    chi_mean = float(line.split("= ")[1].split(" for mean")[0])
    chi_width = float(line.split("and ")[1].split(" for width")[0])
    # two samples of one band over 23 filled bins: chi^2 ~ 23 +- 7, divided by DoF - 1 = 95 as the reference does; the
    # widths of chi^2 fits weighted by the observed contents scatter by more than their parabolic errors (as ROOT's do)
    assert 0.08 < chi_mean < 0.5 and 0.08 < chi_width < 0.9
    ta = table(out)
    both = sel & (ta[:, 4] > 0)
    dof = ana.numBins - 2
    exp_mean = (((t[both, 2] - ta[both, 2]) ** 2) / (t[both, 3] ** 2 + ta[both, 3] ** 2)).sum() / (dof - 1)
    assert chi_mean == pytest.approx(exp_mean, rel=2e-3)
    # one input is not enough for mode 1; unknown modes are refused; the binning must match
    assert run(tool, str(tmp_path / "a.txt"), env={"NB200_MODE": "1"})[0] == 1
    assert run(tool, str(tmp_path / "a.txt"), str(tmp_path / "band.txt"), env={"NB200_MODE": "2"})[0] == 1
    with open(tmp_path / "shifted.txt", "w") as f:
        for r in t:
            f.write("%f\t%f\t%f\t%f\t%f\t%f\n" % (r[0] + 0.3, r[1], r[2], 0.01, 0.1, 0.01))
    rc, _, err = run(tool, str(tmp_path / "a.txt"), str(tmp_path / "shifted.txt"), env={"NB200_MODE": "1"})
    assert rc == 1 and "Binning doesn't match for GoF calculation" in err
    assert run(tool)[0] == 1


def test_skew_gaussian_mode(tool, tmp_path):
    """skewness = 1, the reference's default (analysis.hh:67): every S1 slice re-binned on mean +- 3 sigma and fitted with
    rootNEST's "skewband" (:1319-1330, :1361-1440); leakage from the skew-normal CDF (:726-733)"""
    ana = capi.analysis(0)
    e1, e2 = write_execnest_file(tmp_path / "er.txt", 6, 400000, 2.45, -0.004, 0.16, alpha=2.0, s1max=22.0)
    write_execnest_file(tmp_path / "nr.txt", 7, 100000, 2.2, -0.002, 0.12, s1max=22.0)
    rc, out, err = run(tool, str(tmp_path / "er.txt"), str(tmp_path / "nr.txt"), env={"NB200_SKEWNESS": "1"})
    assert rc == 0, err
    nr_t = table(out, "Calculating band for first dataset")
    er_t = table(out, "Calculating band for second dataset")
    w1, w2 = windowed(ana, e1, e2)
    width = (ana.maxS1 - ana.minS1) / ana.numBins
    ok = (w1 >= 0) & (w2 >= 0)
    lines = err.splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("#evts\tBinCen"))
    rows = np.array([[float(v) for v in ln.split("\t")[:12]] for ln in lines[k + 1:k + 1 + ana.numBins]])
    checked = 0
    for i in range(2, 18):
        lo = ana.minS1 + i * width
        m = ok & (w1 > lo) & (w1 <= lo + width)
        y = np.log10(w2[m] / w1[m])
        mu, sd = y.mean(), y.std(ddof=1)
        a, b = mu - 3 * sd, mu + 3 * sd                     # no wings: the axis stays below 2 / 3 at its ends? checked here
        if a > 2.0 or b > 3.0:
            a, b = a - 0.5, b + 0.5
        h = np.histogram(y, bins=ana.logBins, range=(a, b))[0]
        par, perr, chi = capi.hist_fit(1, h, a, b, (m.sum() * (b - a) / ana.logBins, er_t[i, 2] - 0.1, 0.17, 1.5))
        d = par[3] / np.sqrt(1 + par[3] ** 2)
        if perr[3] > 0.6:
            continue                                         # a slice whose skew the fit cannot pin down: start-dependent
        assert er_t[i, 2] == pytest.approx(par[1] + par[2] * d * np.sqrt(2 / np.pi), abs=2e-3)          # band mean
        assert er_t[i, 4] == pytest.approx(par[2] * np.sqrt(1 - 2 * d * d / np.pi), abs=2e-3)           # band sigma
        assert er_t[i, 6] == pytest.approx(par[3], abs=0.05) and er_t[i, 7] == pytest.approx(perr[3], rel=0.05)
        assert er_t[i, 8] == pytest.approx(chi, rel=0.02)
        assert 0.8 < er_t[i, 6] < 3.0                # the skew that was written (2), as 30 bins weighted by content see it
        # leakage column = skew-normal CDF of the fitted ER slice at the NR mean
        leak, _, _ = capi.band_leakage_skew([par[1]], [par[2]], [perr[2]], [par[3]], [nr_t[i, 2]])
        assert rows[i, 4] == pytest.approx(leak[0], rel=0.05, abs=1e-7)
        checked += 1
    assert checked >= 10
    # fewer than 20000 events: the reference's warning, and Gaussian fits (:1037-1042)
    write_execnest_file(tmp_path / "few.txt", 8, 5000, 2.45, -0.004, 0.16)
    rc, out, err = run(tool, str(tmp_path / "few.txt"), env={"NB200_SKEWNESS": "1"})
    assert rc == 0 and "WARNING: Not enough stats (at least 10^5 events) for skew fits so doing Gaussian" in err
    assert np.all(table(out)[:, 6] == 0)


def test_skewness_2_table_and_propagated_errors(tool, tmp_path):
    """skewness = 2: the 14-column table (rootNEST.cpp:1168-1179) with the fit's covariances, mean / stddev errors through
    them (:1419-1448) and the leakage error propagated from them (:774-845)"""
    ana = capi.analysis(0)
    write_execnest_file(tmp_path / "er.txt", 6, 400000, 2.45, -0.004, 0.16, alpha=2.0, s1max=22.0)
    write_execnest_file(tmp_path / "nr.txt", 7, 100000, 2.2, -0.002, 0.12, s1max=22.0)
    files = (str(tmp_path / "er.txt"), str(tmp_path / "nr.txt"))
    rc1, out1, err1 = run(tool, *files, env={"NB200_SKEWNESS": "1"})
    rc2, out2, err2 = run(tool, *files, env={"NB200_SKEWNESS": "2"})
    assert rc1 == 0 and rc2 == 0
    t1 = table(out1, "Calculating band for second dataset")
    lines = out2.splitlines()
    lines = lines[lines.index("Calculating band for second dataset"):]
    k = next(i for i, ln in enumerate(lines) if ln.startswith("Bin Center\tBand Xi"))
    t2 = np.array([[float(v) for v in ln.split("\t")] for ln in lines[k + 1:k + 1 + ana.numBins]])
    assert t2.shape == (ana.numBins, 14)
    nr2 = out2.splitlines()
    nr2 = nr2[nr2.index("Calculating band for first dataset"):]
    k = next(i for i, ln in enumerate(nr2) if ln.startswith("Bin Center\tBand Xi"))
    tn = np.array([[float(v) for v in ln.split("\t")] for ln in nr2[k + 1:k + 1 + ana.numBins]])
    el = err2.splitlines()
    k = next(i for i, ln in enumerate(el) if ln.startswith("#evts\tBinCen"))
    rows = np.array([[float(v) for v in ln.split("\t")[:12]] for ln in el[k + 1:k + 1 + ana.numBins]])
    checked = 0
    for i in range(2, 18):
        xi, xe, om, oe, al, ae, cxo, cxa, coa, mean, me, sd, se = t2[i, 1:]
        if ae > 0.6:
            continue
        # the same fit as skewness = 1: mean, sigma, alpha
        assert mean == pytest.approx(t1[i, 2], abs=2e-6) and sd == pytest.approx(t1[i, 4], abs=2e-6) and al == pytest.approx(t1[i, 6], abs=2e-6)
        d = al / np.sqrt(1 + al * al)
        assert mean == pytest.approx(xi + om * d * np.sqrt(2 / np.pi), abs=3e-6)
        # :1419-1431 with the reference's own coefficients
        a, b, c = om * np.sqrt(2 / np.pi) * d, np.sqrt(2 / np.pi) * d, om * np.sqrt(2 / np.pi) * (1 + al * al) ** -1.5
        var = a * a * xe * xe + b * b * oe * oe + c * c * ae * ae + 2 * a * b * cxo + 2 * b * c * coa + 2 * a * c * cxa
        assert me == pytest.approx(np.sqrt(var), rel=0.02, abs=2e-6)
        # |cov| <= product of the errors; the leakage error bars are symmetric and equal the library's propagation
        assert abs(cxo) <= xe * oe * 1.001 + 1e-6 and abs(cxa) <= xe * ae * 1.001 + 1e-6 and abs(coa) <= oe * ae * 1.001 + 1e-6
        assert rows[i, 5] == rows[i, 6]
        leak, perr = capi.band_leakage_skew_cov([[xi, xe, om, oe, al, ae, cxo, cxa, coa]], [tn[i, 10]], [tn[i, 11]])
        assert rows[i, 4] == pytest.approx(leak[0], rel=0.02, abs=1e-7) and rows[i, 5] == pytest.approx(perr[0], rel=0.05, abs=1e-7)
        checked += 1
    assert checked >= 10


def test_mono_energetic_peak_table(tool, tmp_path):
    """numBins = 1 (GetFile's peak branch, rootNEST.cpp:1044-1137): Gaussian fits of the S1, S2 and energy histograms,
    means, resolutions in per cent, their errors, chi^2 / DOF -- in the reference's table layout"""
    rng = np.random.default_rng(21)
    n = 40000
    E, s1, s2 = rng.normal(10.0, 0.9, n), rng.normal(60.0, 8.0, n), rng.normal(2600.0, 450.0, n)
    with open(tmp_path / "mono.txt", "w") as f:
        f.write("E_recon [keV]\tfield [V/cm]\ttDrift [us]\tX,Y,Z [mm]\tNph\tNe-\tS1 [PE or phe]\tS1_3Dcor [phd]\t"
                "spikeC(NON-INT)\tNe-Extr\tS2_rawArea [PE]\tS2_3Dcorr [phd]\n")
        for e, a, b in zip(E, s1, s2):
            f.write("%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d\t%.6f\t%.6f\t%.6f\t%d\t%.6f\t%.6f\n" %
                    (e, 180.0, 100.0, 10, -20, 250, 500, 250, a * 1.1, a * 0.97, a, 150, b * 0.8, b))
    rc, out, err = run(tool, str(tmp_path / "mono.txt"), env={"NB200_NUMBINS": "1"})
    assert rc == 0, err
    lines = out.splitlines()
    assert lines[0] == "S1 Mean\t\tS1 Res [%]\tS2 Mean\t\tS2 Res [%]\tEc Mean\t\tEc Res[%]"
    vals = [float(v) for v in lines[1].split("\t") if v.strip()]
    assert lines[2].split("\t\t") == ["+/-"] * 6
    errs = [float(v) for v in lines[3].split("\t")]
    assert lines[4] == "X^2 / DOF\t\t\tX^2 / DOF\t\t\tX^2 / DOF" and lines[5].endswith("\tEnd Table")
    chis = [float(v) for v in lines[5].replace("\tEnd Table", "").split("\t\t\t")]
    for k, (mu, sg) in enumerate(((60.0, 8.0), (2600.0, 450.0), (10.0, 0.9))):
        mean, res, emean, eres = vals[2 * k], vals[2 * k + 1], errs[2 * k], errs[2 * k + 1]
        assert abs(mean - mu) < 5 * emean + 0.002 * mu and abs(res - sg / mu * 100) < 5 * eres + 0.1
        assert emean == pytest.approx(sg / np.sqrt(n), rel=0.3) and eres == pytest.approx(sg / np.sqrt(2 * n) / mu * 100, rel=0.3)
        assert 0.6 < chis[k] < 1.6
    # mode 1 with numBins = 1 takes the same branch (:595-598)
    rc, out1, _ = run(tool, str(tmp_path / "mono.txt"), str(tmp_path / "mono.txt"), env={"NB200_NUMBINS": "1", "NB200_MODE": "1"})
    assert rc == 0 and out1 == out
    if os.path.exists(refprobe.REF_BIN_LUX):       # and the reference's own mono-energetic output goes through
        with open(tmp_path / "ref.txt", "w") as f:
            subprocess.run([refprobe.REF_BIN_LUX, "20000", "gamma", "10", "10", "180", "-1", "1"], stdout=f,
                           stderr=subprocess.DEVNULL, check=True, timeout=300)
        rc, out, err = run(tool, str(tmp_path / "ref.txt"), env={"NB200_NUMBINS": "1"})
        v = [float(x) for x in out.splitlines()[1].split("\t") if x.strip()]
        assert rc == 0 and 9.8 < v[4] < 10.2 and 5.0 < v[5] < 15.0 and 40 < v[0] < 80 and 1500 < v[2] < 4000


def test_on_the_reference_binary_output(tool, tmp_path):
    """the unmodified reference's own stdout goes through: LUX_Run03 ER vs NR discrimination around 99.5 %"""
    if not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("oracle/_ref/execNEST_ref_lux not built")
    for name, args in (("er.txt", ["60000", "beta", "0", "20", "180", "-1", "1"]),
                       ("nr.txt", ["40000", "NR", "0", "60", "180", "-1", "2"])):
        with open(tmp_path / name, "w") as f:
            subprocess.run([refprobe.REF_BIN_LUX] + args, stdout=f, stderr=subprocess.DEVNULL, check=True, timeout=300)
    # with the reference's defaults: skew-Gaussian fits, Woods curve through the NR means
    rc, out, err = run(tool, str(tmp_path / "er.txt"), str(tmp_path / "nr.txt"), env={"NB200_SKEWNESS": "1", "NB200_NRBANDCENTER": "3"})
    assert rc == 0, err
    overall = [ln for ln in err.splitlines() if ln.startswith("OVERALL DISCRIMINATION (ER)")][0]
    leak = float(overall.split("Leakage Fraction = ")[1])
    assert 1e-3 < leak < 1.5e-2
    gauss = [ln for ln in err.splitlines() if ln.startswith("OVERALL DISCRIMINATION      ")][0]
    assert 1e-3 < float(gauss.split("Leakage Fraction = ")[1]) < 1.5e-2


def test_goodness_of_fit_against_the_lux_tritium_band(tool, tmp_path):
    """the workflow the mode exists for (reference README.md:405-420): a simulated CH3T band against the LUX Run03 tritium
    band the reference ships (examples/LUXRun03_CH3TBand_SpikeFull.txt, same 98 S1 bins as analysis.hh).  Simulation =
    the unmodified reference binary here (no GPU in this container); on a GPU box execNEST_b200 takes its place."""
    data = "/root/reference/examples/LUXRun03_CH3TBand_SpikeFull.txt"
    if not (os.path.exists(data) and os.path.exists(refprobe.REF_BIN_LUX)):
        pytest.skip("reference tree / oracle/_ref not present")
    with open(tmp_path / "ch3t.txt", "w") as f:
        subprocess.run([refprobe.REF_BIN_LUX, "100000", "CH3T", "0", "18.6", "180", "-1", "1"], stdout=f,
                       stderr=subprocess.DEVNULL, check=True, timeout=600)
    for skew in ("0", "1"):
        rc, out, err = run(tool, str(tmp_path / "ch3t.txt"), data, env={"NB200_MODE": "1", "NB200_SKEWNESS": skew})
        assert rc == 0, err
        line = [ln for ln in out.splitlines() if ln.startswith("%-errors of the form")][0]
        pm, pw = [float(v) for v in re.findall(r"-?\d+\.\d+(?:e[-+]\d+)?", line.split("are ")[1])[:2]]
        # NEST's LUX model sits 0.7 % under the measured band means (2.6431 simulated vs 2.6644 measured in the
        # first bin) with widths within a few per cent
        assert -2.0 < pm < 0.5 and abs(pw) < 6.0, (skew, line)
        chi = [ln for ln in out.splitlines() if ln.startswith("The reduced CHI^2 = ")][0]
        assert float(chi.split("= ")[1].split(" for mean")[0]) < 40.0
