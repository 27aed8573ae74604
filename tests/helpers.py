"""Shared helpers for the parity tests: run the GPU chain and lay its columns out like the
reference probe's (oracle/ref_probe.cpp) so that both sides are compared column by column."""
import numpy as np
from scipy import stats

import refprobe
from nest_b200 import capi

RHO_LUX = 2.8886336386968878


def gpu_chain(ctx, seed, n, species, src_kind, eMin, eMax, inField=180.0, fixed_pos=0, xyz=(0, 0, 0),
              er_model=-1, multFact=1.0, model_which=1, skewER=None, ana=None, det=None, band=None,
              first_event=0, par=(), wimp=None):
    det = det or capi.detector("LUX_Run03")
    rc = capi.prepare(det, inField, (ana.z_step if ana else 0.0))
    mo = capi.model(model_which)
    mo.er_model = er_model
    mo.multFact = multFact
    if skewER is not None:
        mo.SkewnessER = skewER
    # massNum / atomNum as the CLI sets them (execNEST.cpp:438-439, 496-500, 670-671, 1052)
    if species == capi.ion:
        mo.massNum, mo.atomNum = 4.0, 2.0
    elif species == capi.Kr83m:
        mo.massNum, mo.atomNum = eMax, 0.0
    elif species >= 6:
        mo.massNum, mo.atomNum = det.molarMass, 0.0
    ana = ana or capi.analysis(0)
    src = capi.source(src_kind, species, eMin, eMax, inField, fixed_pos, xyz, par)
    if wimp is not None:  # capi.WimpTable (TestSpectra::WIMP_spectrum_prep)
        wimp.attach(src)
    cols = ctx.run(det, mo, ana, src, rc, seed, n, first_event=first_event, band=band)
    out = np.zeros((n, refprobe.SIGE + 1))
    C = refprobe.COL
    out[:, C["keV_true"]] = cols["energy_kev"]
    out[:, C["keV_out"]] = cols["signalE"]
    out[:, C["field"]] = cols["field"]
    out[:, C["dt"]] = cols["driftTime"]
    out[:, C["x"]] = cols["x_mm"]
    out[:, C["y"]] = cols["y_mm"]
    out[:, C["z"]] = cols["z_mm"]
    out[:, C["sx"]] = cols["smear_x"]
    out[:, C["sy"]] = cols["smear_y"]
    out[:, C["Nph"]] = cols["photons"]
    out[:, C["Ne"]] = cols["electrons"]
    out[:, C["Ni"]] = cols["ions"]
    out[:, C["Nex"]] = cols["excitons"]
    out[:, C["Yph"]] = cols["Nph_mean"]
    out[:, C["Yne"]] = cols["Ne_mean"]
    for k in range(9):
        out[:, refprobe.S1_0 + k] = cols["s1_%d" % k]
        out[:, refprobe.S2_0 + k] = cols["s2_%d" % k]
    out[:, refprobe.SIG1] = cols["signal1"]
    out[:, refprobe.SIG2] = cols["signal2"]
    out[:, refprobe.SIGE] = cols["signalE"]
    return out, cols


def ref_chain(probe, seed, n, species, src_kind, eMin, eMax, inField=180.0, fixed_pos=0, xyz=(0, 0, 0),
              er_model=-1, multFact=1.0, model_which=1, skewER=None, ana=None):
    mo = capi.model(model_which)
    sk = mo.SkewnessER if skewER is None else skewER
    # the reference's GetYieldGamma branch takes -multFact (execNEST.cpp:1044-1045)
    mf = -multFact if er_model == 1 else multFact
    return probe.chain(seed, n, species, src_kind, eMin, eMax, inField, fixed_pos, list(xyz), er_model, mf,
                       list(mo.NRYieldsParam), list(mo.ERYieldsParam), list(mo.NRERWidthsParam)[:mo.n_widths],
                       sk, 0, ana=ana or capi.analysis(0))


COLUMN_NAMES = (refprobe.COLS + ["s1_%d" % k for k in range(9)] + ["s2_%d" % k for k in range(9)] +
                ["signal1", "signal2", "signalE"])


def _sig(v, digits=9):
    with np.errstate(divide="ignore", invalid="ignore"):
        mag = np.where(v != 0, np.floor(np.log10(np.abs(v))), 0.0)
    f = 10.0 ** (digits - 1 - mag)
    return np.round(v * f) / f


def ks_report(a, b, skip=()):
    """two-sample KS p-value per column (columns that are constant on both sides must be equal)"""
    res = {}
    for j, name in enumerate(COLUMN_NAMES):
        if name in skip:
            continue
        x, y = a[:, j], b[:, j]
        if name in ("Yph", "Yne"):
            # deterministic means: the two libm's differ in the last ulp, which would split the
            # atoms of a discrete distribution (9 Xe isotopes) -- compare at 1e-9 relative
            x, y = _sig(x), _sig(y)
        if np.all(x == x[0]) and np.all(y == y[0]):
            res[name] = 1.0 if np.isclose(x[0], y[0], rtol=1e-12, atol=0) or x[0] == y[0] else 0.0
            continue
        res[name] = stats.ks_2samp(x, y).pvalue
    return res


def mean_z(a, b):
    """z-score of the difference of means per column"""
    out = {}
    for j, name in enumerate(COLUMN_NAMES):
        x, y = a[:, j], b[:, j]
        se = np.sqrt(x.var() / x.size + y.var() / y.size)
        out[name] = 0.0 if se == 0 else float((x.mean() - y.mean()) / se)
    return out


def band_chi2(b1, b2, free_param=2):
    """reduced chi^2 of band means and widths, examples/rootNEST.cpp:619-643"""
    ok = np.isfinite(b1).all(axis=1) & np.isfinite(b2).all(axis=1) & (b1[:, 4] > 0) & (b2[:, 4] > 0)
    b1, b2 = b1[ok], b2[ok]
    dof = len(b1) - free_param
    e2 = b1[:, 4] ** 2 + b2[:, 4] ** 2
    chi_mean = np.sum((b2[:, 2] - b1[:, 2]) ** 2 / e2) / (dof - 1)
    # error of a standard deviation ~ sigma/sqrt(2N) = (sigma/sqrt(N))/sqrt(2)
    chi_width = np.sum((b2[:, 3] - b1[:, 3]) ** 2 / (e2 / 2.0)) / (dof - 1)
    return chi_mean, chi_width, dof
