"""The analysis.hh switches other than the two shipped settings (usePD = 2 / useS2 = 0 with LUX_Run03, usePD = 1 /
useS2 = 1 in analysis_G2.hh): usePD 0 / 1 / 2 (pulse areas in phe, phd, spike counts: execNEST.cpp:1121-1158 signal
selection, :1170-1250 energy reconstruction), useS2 0 / 1 / 2 (band axes, GetBand :1508-1621), MCtruthE, MCtruthPos.
The chain's selection, reconstruction and band on the GPU against the CPU restatement (oracle port) event column by
event column; the restatement and the GPU command line are pinned to the reference itself in these modes by
test_cli_analysis_modes_match_the_reference_binary (the reference binary with its analysis globals set at start-up)."""
import os
import re
import subprocess

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import band_chi2, gpu_chain, ks_report, ref_chain
from nest_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODES = [  # usePD, useS2, MCtruthE, MCtruthPos, minS1, maxS1, minS2, maxS2
    (0, 0, 0, 0, 2.0, 165.0, 80.0, 2.0e4),
    (1, 0, 0, 0, 1.5, 120.0, 60.0, 1.5e4),
    (0, 1, 0, 0, 2.0, 165.0, 80.0, 2.0e4),
    (2, 1, 1, 0, 1.5, 99.5, 42.0, 1.0e4),
    (1, 2, 0, 1, 1.5, 120.0, 60.0, 1.5e4),
    (2, 2, 1, 1, 1.5, 99.5, 42.0, 1.0e4),
]


def _ana(mode):
    usePD, useS2, tE, tP, a, b, c, d = mode
    ana = capi.analysis(0)
    ana.usePD, ana.useS2, ana.MCtruthE, ana.MCtruthPos = usePD, useS2, tE, tP
    ana.minS1, ana.maxS1, ana.minS2, ana.maxS2 = a, b, c, d
    if useS2 == 1:
        ana.logMin, ana.logMax = 1.5, 4.5     # y = log10(S2)
    return ana


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("species,er_model,emin,emax", [(capi.beta, 0, 0.1, 20.0), (capi.NR, -1, 0.5, 60.0)])
def test_analysis_modes_against_the_restatement(ctx, mode, species, er_model, emin, emax):
    orc = refprobe.oracle()
    if orc is None:
        pytest.skip("oracle/liboracle.so not built")
    ana = _ana(mode)
    n = 40000
    band = capi.Band(ana)
    g, _ = gpu_chain(ctx, 11, n, species, capi.SRC_FLAT, emin, emax, er_model=er_model, ana=ana, band=band)
    r = ref_chain(orc, 12, n, species, capi.SRC_FLAT, emin, emax, er_model=er_model, ana=ana)
    # (helpers.gpu_chain fills "keV_out" with signalE, the probes with the printed energy: the same number unless
    # MCtruthE keeps the true energy there while signalE of a sub-threshold event is 0 -- signalE has its own column)
    ks = ks_report(g, r, skip=("keV_out",) if mode[2] else ())
    bad = {k: v for k, v in ks.items() if v < 1e-4}
    assert not bad, (mode, bad)
    # the selected signals must not be trivial in these settings
    assert (g[:, refprobe.SIG1] > 0).mean() > 0.2 and (g[:, refprobe.SIG2] > 0).mean() > 0.2
    # band: the GPU accumulator against GetBand (restated) on the GPU's own signals -- same events, so equal to rounding
    # -- and against GetBand on the restatement's events statistically
    gb = band.finalize()
    # execNEST.cpp:1424-1427: useS2 == 2 hands GetBand the two signal vectors the other way round
    a1, a2 = (refprobe.SIG2, refprobe.SIG1) if ana.useS2 == 2 else (refprobe.SIG1, refprobe.SIG2)
    bs = orc.getband(g[:, a1], g[:, a2], ana=ana)
    assert len(bs) == len(gb)
    # (bins of a few events are left out: their width is the difference of two fixed-point sums of nearly equal size)
    ok = np.isfinite(bs).all(axis=1) & np.isfinite(gb).all(axis=1) & (band.accepted > 30)
    assert ok.sum() >= 5, (mode, ok.sum())
    assert np.allclose(gb[ok][:, :6], bs[ok][:, :6], rtol=1e-5, atol=1e-7), (mode, np.abs(gb[ok][:, :6] / bs[ok][:, :6] - 1).max(axis=0))
    rb = orc.getband(r[:, a1], r[:, a2], ana=ana)
    pop = ok & np.isfinite(rb).all(axis=1) & (band.accepted > 100)
    if pop.sum() > 5:
        cm, cw, dof = band_chi2(gb[pop], rb[pop])
        assert cm < 2.5 and cw < 2.5, (mode, cm, cw, dof)


CLI = os.path.join(ROOT, "nest_b200", "host", "execNEST_b200")


def _env(mode, prefix):
    usePD, useS2, tE, tP, a, b, c, d = mode
    e = dict(os.environ)
    e.update({prefix + "USEPD": str(usePD), prefix + "USES2": str(useS2), prefix + "MCTRUTHE": str(tE),
              prefix + "MCTRUTHPOS": str(tP), prefix + "MINS1": repr(a), prefix + "MAXS1": repr(b),
              prefix + "MINS2": repr(c), prefix + "MAXS2": repr(d)})
    if useS2 == 1:
        e.update({prefix + "LOGMIN": "1.5", prefix + "LOGMAX": "4.5"})
    return e


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("args", [("30000", "ER", "0.1", "20", "180", "-1", "1"), ("30000", "NR", "0.5", "60", "180", "-1", "2")])
def test_cli_analysis_modes_match_the_reference_binary(mode, args):
    """the command line in the same modes against the reference ITSELF: oracle/_ref/execNEST_ref_lux_env is the
    unmodified src/execNEST.cpp whose analysis.hh globals are set from the environment before main() (oracle/
    shim_lux_env); same argv, same header lines, event columns by KS, band table by rootNEST's chi^2"""
    from test_gpu_cli import band_of, rows_of
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX_ENV):
        pytest.skip("execNEST_b200 / execNEST_ref_lux_env not built")
    rg = subprocess.run([CLI] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=_env(mode, "NB200_"))
    rr = subprocess.run([refprobe.REF_BIN_LUX_ENV] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                        env=_env(mode, "NEST_REF_"))
    assert rg.returncode == 0 and rr.returncode == 0, (rg.stderr[-400:], rr.stderr[-400:])
    hg = [ln for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert len(hg) == len(hr)
    for a, b in zip(hg, hr):
        if a.startswith("g1 = "):
            assert a.split("SE_mean")[0] == b.split("SE_mean")[0]
        else:
            assert a == b, (mode, a, b)          # the column titles change with usePD / MCtruthE / MCtruthPos
    tg, tr = _stderr_text(rg.stderr), _stderr_text(rr.stderr)
    assert tg == tr, "%s %s\nonly here:\n  %s\nonly in the reference:\n  %s" % (mode, args, "\n  ".join(sorted(tg - tr)), "\n  ".join(sorted(tr - tg)))
    g, r = rows_of(rg.stdout), rows_of(rr.stdout)
    assert g.shape == r.shape and g.shape[1] == 14 and g.shape[0] > 0.5 * int(args[0]), (g.shape, r.shape)
    names = ["keV", "field", "dt", "x", "y", "z", "Nph", "Ne", "S1", "S1c", "spikeC", "Nee", "S2", "S2c"]
    for j, nme in enumerate(names):
        if np.all(g[:, j] == g[0, j]):
            assert np.all(r[:, j] == g[0, j]), nme
            continue
        p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
        assert p > 1e-4, (mode, nme, p, g[:, j].mean(), r[:, j].mean())
    bg, br = band_of(rg.stderr), band_of(rr.stderr)
    assert bg.shape == br.shape == (98, 7), (bg.shape, br.shape)
    assert np.array_equal(bg[:, 0], br[:, 0])
    ok = np.isfinite(bg).all(axis=1) & np.isfinite(br).all(axis=1) & (bg[:, 4] > 0) & (br[:, 4] > 0)
    if ok.sum() > 8:
        cm, cw, dof = band_chi2(bg[ok], br[ok])
        assert cm < 2.5 and cw < 2.5, (mode, cm, cw, dof)


def _numeric_rows(out):
    rows = []
    for ln in out.splitlines():
        if re.match(r"^-?[0-9]", ln):
            try:
                rows.append([float(v) for v in re.split(r"[\t,]\s*", ln.strip()) if v != ""])
            except ValueError:
                pass
    width = max((len(r) for r in rows), default=0)
    return np.array([r for r in rows if len(r) == width])


@pytest.mark.parametrize("extra,args", [
    ({"PRINTSUBTHR": "-1"}, ("20000", "NR", "0", "25", "180", "-1", "1")),     # rows inside the S1 / S2 windows only
    ({"PRINTSUBTHR": "0"}, ("20000", "NR", "0", "25", "180", "-1", "1")),      # rows above threshold only
    ({"PRINTSUBTHR": "-1", "USEPD": "0", "MINS1": "3.0", "MAXS1": "60.0"}, ("20000", "ER", "0.1", "12", "180", "-1", "1")),
    ({"S1MODE": "0"}, ("20000", "gamma", "60", "300", "180", "-1", "1")),       # Full S1 also above 100 keV
    ({"S1MODE": "1"}, ("20000", "ER", "0.5", "25", "180", "-1", "1")),          # Parametric S1 also below 100 keV
    ({"S1MODE": "1", "USEPD": "1", "MINS1": "2.0", "MAXS1": "110."}, ("20000", "NR", "1", "60", "180", "-1", "2")),
    ({"NUMBINS": "25", "MAXS1": "76.5"}, ("20000", "ER", "0.1", "20", "180", "-1", "1")),   # another band binning
    ({"VERBOSITY": "2"}, ("8000", "ER", "0.1", "20", "180", "-1", "1")),        # rows carry Z alone (execNEST.cpp:1377-1380)
    ({"VERBOSITY": "0"}, ("8000", "ER", "0.1", "20", "180", "-1", "1")),
    ({"VERBOSITY": "-1"}, ("8000", "ER", "0.1", "20", "180", "-1", "1")),
])
def test_cli_print_switches_match_the_reference_binary(extra, args):
    """PrintSubThr (-1 / 0 / 1: which events get a row, execNEST.cpp:1336-1359) and verbosity (what is printed at all, the
    row format): number of rows, header lines and columns against the reference binary with the same switches"""
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX_ENV):
        pytest.skip("execNEST_b200 / execNEST_ref_lux_env not built")
    eg, er = dict(os.environ), dict(os.environ)
    for k, v in extra.items():
        eg["NB200_" + k] = v
        er["NEST_REF_" + k] = v
    rg = subprocess.run([CLI] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=eg)
    rr = subprocess.run([refprobe.REF_BIN_LUX_ENV] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=er)
    assert rg.returncode == rr.returncode == 0, (rg.stderr[-300:], rr.stderr[-300:])
    hg = [ln for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert [a.split("SE_mean")[0] for a in hg] == [b.split("SE_mean")[0] for b in hr], (extra, hg, hr)
    g, r = _numeric_rows(rg.stdout), _numeric_rows(rr.stdout)
    n = int(args[0])
    assert g.ndim == r.ndim
    if r.size == 0:
        assert g.size == 0
    else:
        assert g.shape[1] == r.shape[1], (extra, g.shape, r.shape)
        # the number of rows is itself a random variable when sub-threshold events are left out
        pr = r.shape[0] / n
        assert abs(g.shape[0] - r.shape[0]) <= 5.0 * np.sqrt(2 * n * max(pr * (1 - pr), 1e-9)) + 1e-9, (extra, g.shape, r.shape)
        for j in range(r.shape[1]):
            if np.all(r[:, j] == r[0, j]) and np.all(g[:, j] == g[0, j]):
                assert r[0, j] == g[0, j], (extra, j)
                continue
            p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
            assert p > 1e-4, (extra, j, p, g[:, j].mean(), r[:, j].mean())
    # stderr: the same kinds of lines (band table present or not, warnings) -- compare the non-numeric ones
    def text(err):
        # (the "Insufficient number of ... events" hint depends on which bins a finite sample leaves empty)
        return set(ln for ln in err.splitlines() if ln.strip() and not re.match(r"^-?[0-9]", ln) and "SE_mean" not in ln and
                   "Insufficient number" not in ln)
    tg, tr = text(rg.stderr), text(rr.stderr)
    assert tg == tr, (extra, "only here:", sorted(tg - tr), "only in the reference:", sorted(tr - tg))
    ng = sum(1 for ln in rg.stderr.splitlines() if re.match(r"^[0-9]", ln))
    nr_ = sum(1 for ln in rr.stderr.splitlines() if re.match(r"^[0-9]", ln))
    assert ng == nr_, (extra, ng, nr_)
    from test_gpu_cli import band_of
    bg, br = band_of(rg.stderr), band_of(rr.stderr)
    assert bg.shape == br.shape
    if bg.size:
        assert np.array_equal(bg[:, 0], br[:, 0])
        ok = np.isfinite(bg).all(axis=1) & np.isfinite(br).all(axis=1) & (bg[:, 4] > 0) & (br[:, 4] > 0)
        if ok.sum() > 8:
            cm, cw, dof = band_chi2(bg[ok], br[ok])
            assert cm < 2.5 and cw < 2.5, (extra, cm, cw, dof)


@pytest.mark.parametrize("args", [
    ("20000", "ER", "1", "15", "180", "50,-30,200", "1"),        # fixed vertex
    ("20000", "NR", "2", "40", "180", "50,-30,-1", "2"),          # fixed x, y; z drawn (execNEST.cpp:846-848)
    ("20000", "ER", "1", "15", "180", "-999,-999,150", "3"),      # x, y drawn; fixed z (:849-854)
    ("20000", "ER", "1", "15", "-1", "-1", "4"),                  # FitEF field at a random vertex (:857-861)
    ("20000", "NR", "2", "40", "-1", "30,40,-1", "5"),
    ("20000", "NR", "1", "50", "180", "-1", "-2"),                # negative seed: rows carry the mean yields (:1363-1376)
    ("20000", "beta", "1", "15", "180", "-1", "-5"),
])
def test_cli_position_field_and_seed_modes_match_the_reference_binary(args):
    """the command line's position argument in its three forms, the FitEF field mode and the negative-seed mode (mean
    yields printed in place of the quanta) against the unmodified reference binary: same header, rows column by column"""
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("execNEST_b200 / execNEST_ref_lux not built")
    rg = subprocess.run([CLI] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rr = subprocess.run([refprobe.REF_BIN_LUX] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert rg.returncode == rr.returncode == 0, (rg.stderr[-300:], rr.stderr[-300:])
    hg = [ln.split("SE_mean")[0] for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln.split("SE_mean")[0] for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert hg == hr, (args, hg, hr)
    tg, tr = _stderr_text(rg.stderr), _stderr_text(rr.stderr)
    assert tg == tr, "%s\nonly here:\n  %s\nonly in the reference:\n  %s" % (args, "\n  ".join(sorted(tg - tr)), "\n  ".join(sorted(tr - tg)))
    g, r = _numeric_rows(rg.stdout), _numeric_rows(rr.stdout)
    assert g.shape == r.shape and g.shape[0] == int(args[0]), (args, g.shape, r.shape)
    for j in range(r.shape[1]):
        if np.all(r[:, j] == r[0, j]) or np.all(g[:, j] == g[0, j]):
            assert np.all(g[:, j] == r[0, j]) and np.all(r[:, j] == r[0, j]), (args, j, g[0, j], r[0, j])
            continue
        p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
        assert p > 1e-4, (args, j, p, g[:, j].mean(), r[:, j].mean())


def test_cli_outside_the_detector_fails_like_the_reference():
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("execNEST_b200 / execNEST_ref_lux not built")
    for args in [("100", "ER", "1", "15", "180", "400,0,100", "1"),        # beyond radmax (:835-840)
                 ("100", "ER", "1", "15", "180", "0,0,100"),               # fine (no seed)
                 ("100", "ER", "1", "15", "180", "0;0;100", "1")]:         # not a position
        rg = subprocess.run([CLI] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        rr = subprocess.run([refprobe.REF_BIN_LUX] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert (rg.returncode != 0) == (rr.returncode != 0), (args, rg.returncode, rr.returncode, rg.stderr[-200:], rr.stderr[-200:])


def _stderr_text(err):
    """the non-numeric lines a run leaves on stderr (prompts, notices, warnings), without the ones that depend on the
    sample (which band bins a finite run leaves empty) or carry a Monte-Carlo number"""
    out = set()
    for ln in err.splitlines():
        t = ln.strip()
        if not t or re.match(r"^-?[0-9]", t) or "SE_mean" in t or "Insufficient number" in t:
            continue
        out.add(t)
    return out


@pytest.mark.parametrize("args,stdin,expected", [
    (("15000", "D-D", "0", "80", "180", "-1", "1"), None, 15000),
    (("15000", "AmBe", "0", "200", "180", "-1", "1"), None, 15000),
    (("15000", "252Cf", "0", "200", "180", "-1", "1"), None, 15000),
    (("15000", "14C", "0", "156", "180", "-1", "1"), None, 15000),
    (("15000", "x-ray", "5", "40", "180", "-1", "1"), None, 15000),
    (("15000", "Compton", "0.5", "30", "180", "-1", "1"), None, 15000),
    (("15000", "tritium", "0", "18.6", "180", "-1", "1"), None, 15000),
    (("15000", "-1", "1", "50", "180", "-1", "1"), None, 15000),                 # "-1": the default type, NR (:472-473)
    (("5000000", "8B", "0", "4", "180", "-1", "1"), None, 0.0026 * 5e6),          # Poisson-fluctuated counts (:485-488)
    (("10000000", "pp-solar", "0", "250", "180", "-1", "1"), None, 0.0011794 * 1e7),
    (("60000000000", "atm_nu", "0", "100", "180", "-1", "1"), None, 1.5764e-7 * 6e10),
    (("10000", "alpha", "4000", "6000", "180", "-1", "1"), None, 10000),
    (("10000", "83mKr", "32.1", "300", "180", "-1", "1"), None, 10000),          # the 32.1 keV line alone, times <= 300 ns
    (("10000", "Kr83", "41.5", "1000", "-1", "-1", "1"), None, 10000),           # both lines merged, FitEF field
    (("10000", "nucleus", "10", "200", "180", "-1", "1"), "82\n206\n", 10000),   # A, Z from stdin (:496-505)
    (("10000", "ion", "5", "60", "180", "-1", "1"), "54\n131\n", 10000),         # Z = 54: plain NR (:506)
])
def test_cli_types_match_the_reference_binary(args, stdin, expected):
    """the type argument in its aliases, the Poisson-fluctuated event counts of the rate-normalised sources and the ion's
    A, Z read from stdin: row count and every column against the unmodified reference binary"""
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("execNEST_b200 / execNEST_ref_lux not built")
    rg = subprocess.run([CLI] + list(args), input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rr = subprocess.run([refprobe.REF_BIN_LUX] + list(args), input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert rg.returncode == rr.returncode == 0, (args, rg.stderr[-300:], rr.stderr[-300:])
    hg = [ln.split("SE_mean")[0] for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln.split("SE_mean")[0] for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert hg == hr, (args, hg, hr)
    tg, tr = _stderr_text(rg.stderr), _stderr_text(rr.stderr)
    assert tg == tr, "%s\nonly here:\n  %s\nonly in the reference:\n  %s" % (args, "\n  ".join(sorted(tg - tr)), "\n  ".join(sorted(tr - tg)))
    g, r = _numeric_rows(rg.stdout), _numeric_rows(rr.stdout)
    assert g.shape[1] == r.shape[1], (args, g.shape, r.shape)
    tol = 5.0 * np.sqrt(expected) if expected != int(args[0]) else 0.0
    assert abs(g.shape[0] - expected) <= tol and abs(r.shape[0] - expected) <= tol, (args, g.shape, r.shape, expected)
    for j in range(r.shape[1]):
        if np.all(r[:, j] == r[0, j]) or np.all(g[:, j] == g[0, j]):
            assert np.all(g[:, j] == r[0, j]) and np.all(r[:, j] == r[0, j]), (args, j, g[0, j], r[0, j])
            continue
        p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
        assert p > 1e-4, (args, j, p, g[:, j].mean(), r[:, j].mean())


DET_DIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("det", ["xenon10", "xenonnt", "lz2022", "g3", "zeplin",
                                 # user-style detectors derived from LUX_Run03 (oracle/shim_variants/LUX_variants.hh): bottom-array
                                 # S2 + 3-fold coincidence; sPEeff < 1 with linear / quadratic / baseline noise; no coincidence
                                 # requirement with a short electron lifetime and weak extraction; another (T, p, W) point
                                 "luxbot3", "luxnoisy", "luxnocoin", "luxwarm"])
@pytest.mark.parametrize("args", [("30000", "NR", "0", "70", "-1", "-1", "1"), ("30000", "beta", "0.1", "25", "-1", "-1", "2"),
                                  ("20000", "NR", "1", "50", "250", "-1", "3")])
def test_every_shipped_detector_header_matches_the_reference_built_with_it(det, args):
    """the other detector headers the reference ships (XENON10, XENONnT SR0, LZ WS2022, the projected G3, ZEPLIN-III),
    each UNMODIFIED, and four user-style detectors derived from the reference's LUX_Run03 that reach the detector-parameter
    branches of GetS1 / GetS2 the shipped ones do not: nest-b200's command line compiled with the header against the VDetector mirror (position maps through
    (r,z) lookup tables, `FitEF` field at -1) against the unmodified src/execNEST.cpp instantiating the same class --
    banner (g1, g2, density, drift speed), rows column by column, band table"""
    from test_gpu_cli import band_of, rows_of
    cli, ref = os.path.join(DET_DIR, "execNEST_b200_" + det), os.path.join(DET_DIR, "execNEST_ref_" + det)
    if not (os.path.exists(cli) and os.path.exists(ref)):
        pytest.skip("oracle/_ref/execNEST_{b200,ref}_%s not built (needs the reference tree at build time)" % det)
    rg = subprocess.run([cli] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    rr = subprocess.run([ref] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert rg.returncode == rr.returncode == 0, (det, rg.stderr[-400:], rr.stderr[-400:])
    hg = [ln.split("SE_mean")[0] for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln.split("SE_mean")[0] for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert hg == hr, (det, hg, hr)
    tg, tr = _stderr_text(rg.stderr), _stderr_text(rr.stderr)
    assert tg == tr, "%s %s\nonly here:\n  %s\nonly in the reference:\n  %s" % (det, args, "\n  ".join(sorted(tg - tr)), "\n  ".join(sorted(tr - tg)))
    g, r = rows_of(rg.stdout), rows_of(rr.stdout)
    n = int(args[0])
    assert g.shape[1] == r.shape[1] == 14 and 0 < r.shape[0] <= n
    assert abs(g.shape[0] - r.shape[0]) < 5 * np.sqrt(2 * max(n - r.shape[0], 25)), (det, g.shape, r.shape)
    names = ["keV", "field", "dt", "x", "y", "z", "Nph", "Ne", "S1", "S1c", "spikeC", "Nee", "S2", "S2c"]
    for j, nme in enumerate(names):
        if np.all(g[:, j] == g[0, j]) or np.all(r[:, j] == r[0, j]):
            assert np.all(r[:, j] == g[0, j]) and np.all(g[:, j] == g[0, j]), (det, nme, g[0, j], r[0, j])
            continue
        p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
        assert p > 1e-4, (det, nme, p, g[:, j].mean(), r[:, j].mean())
    bg, br = band_of(rg.stderr), band_of(rr.stderr)
    assert bg.shape == br.shape and bg.shape[0] > 10
    assert np.array_equal(bg[:, 0], br[:, 0])
    ok = np.isfinite(bg).all(axis=1) & np.isfinite(br).all(axis=1) & (bg[:, 4] > 0) & (br[:, 4] > 0)
    if ok.sum() > 8:
        cm, cw, dof = band_chi2(bg[ok], br[ok])
        assert cm < 2.5 and cw < 2.5, (det, cm, cw, dof)


@pytest.mark.parametrize("args", [
    ("10000", "NR", "1", "50", "0", "-1", "1"),             # zero field: warned about, simulated (S2 vanishes)
    ("10000", "beta", "1", "20", "0.5", "-1", "1"),         # below FIELD_MIN
    ("10000", "NR", "1", "50", "5000", "-1", "1"),          # beyond the drift-speed table's range
    ("10000", "gamma", "200", "1500", "180", "-1", "1"),    # "Energy [keV]" header, %e rows above 1 MeV, Parametric S1 (Hybrid)
    ("10000", "NR", "100", "330", "180", "-1", "1"),        # up to the NR model's validity cap
    ("10000", "beta", "500", "3000", "180", "-1", "1"),
    ("10000", "NR", "0.001", "0.6", "180", "-1", "1"),      # mostly below one quantum
    ("10000", "beta", "0.005", "0.3", "180", "-1", "1"),
])
def test_cli_field_and_energy_regimes_match_the_reference_binary(args):
    """the corners of the field and energy arguments: zero and sub-threshold fields, fields beyond the drift-speed fit,
    MeV energies (header, row format, S1 mode switch), the NR validity cap, energies of a few quanta"""
    if not os.path.exists(CLI) or not os.path.exists(refprobe.REF_BIN_LUX):
        pytest.skip("execNEST_b200 / execNEST_ref_lux not built")
    rg = subprocess.run([CLI] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rr = subprocess.run([refprobe.REF_BIN_LUX] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert rg.returncode == rr.returncode, (args, rg.returncode, rr.returncode, rg.stderr[-300:], rr.stderr[-300:])
    if rr.returncode != 0:
        return
    hg = [ln.split("SE_mean")[0] for ln in rg.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    hr = [ln.split("SE_mean")[0] for ln in rr.stdout.splitlines() if not re.match(r"^-?[0-9]", ln)]
    assert hg == hr, (args, hg, hr)
    tg, tr = _stderr_text(rg.stderr), _stderr_text(rr.stderr)
    assert tg == tr, "%s\nonly here:\n  %s\nonly in the reference:\n  %s" % (args, "\n  ".join(sorted(tg - tr)), "\n  ".join(sorted(tr - tg)))
    g, r = _numeric_rows(rg.stdout), _numeric_rows(rr.stdout)
    assert g.shape == r.shape and g.shape[0] == int(args[0]), (args, g.shape, r.shape)
    for j in range(r.shape[1]):
        if np.all(r[:, j] == r[0, j]) or np.all(g[:, j] == g[0, j]):
            assert np.all(g[:, j] == r[0, j]) and np.all(r[:, j] == r[0, j]), (args, j, g[0, j], r[0, j])
            continue
        p = stats.ks_2samp(g[:, j], r[:, j]).pvalue
        assert p > 1e-4, (args, j, p, g[:, j].mean(), r[:, j].mean())
