"""High-statistics parity: 4e7 GPU events against 4.8e6 events of the unmodified reference (one
process per core, distinct seeds), config 2 (NR 0-100 keV) and config 1 (ER 0.1-20 keV).  Band means
and widths by the reference's chi^2 (examples/rootNEST.cpp:619-643) and 1-D chi^2 on the integer
quanta and corrected pulse areas."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["Nph", "Ne", "s1c", "s2c"]


def _edges(cfg):
    if cfg == "NR":
        return {"Nph": np.arange(0, 1801, 20.), "Ne": np.arange(0, 361, 4.), "s1c": np.arange(0, 241, 3.),
                "s2c": np.arange(0, 6001, 75.)}
    return {"Nph": np.arange(0, 1601, 20.), "Ne": np.arange(0, 801, 10.), "s1c": np.arange(0, 201, 2.5),
            "s2c": np.arange(0, 12001, 150.)}


def _ref_worker(args):
    cfg, seed, n = args
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refprobe
    from helpers import ref_chain
    from nest_b200 import capi, shard
    r = refprobe.reference()
    if cfg == "NR":
        a = ref_chain(r, seed, n, capi.NR, capi.SRC_FLAT, 0.0, 100.0)
    else:
        a = ref_chain(r, seed, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0)
    ana = capi.analysis(0)
    words = shard.band_words_from_signals(ana, 2, a[:, refprobe.SIG1], a[:, refprobe.SIG2])
    ed = _edges(cfg)
    h = {"Nph": np.histogram(a[:, refprobe.COL["Nph"]], ed["Nph"])[0],
         "Ne": np.histogram(a[:, refprobe.COL["Ne"]], ed["Ne"])[0],
         "s1c": np.histogram(np.abs(a[:, refprobe.S1_0 + 5]), ed["s1c"])[0],
         "s2c": np.histogram(np.abs(a[:, refprobe.S2_0 + 7]), ed["s2c"])[0]}
    return words, h


def chi2_hist(a, b):
    """two-sample chi^2 of histograms with different totals; bins with few counts merged away"""
    a, b = a.astype(float), b.astype(float)
    ok = (a + b) > 50
    a, b = a[ok], b[ok]
    A, B = a.sum(), b.sum()
    c = np.sum((np.sqrt(B / A) * a - np.sqrt(A / B) * b) ** 2 / (a + b))
    dof = len(a) - 1
    return c, dof, stats.chi2.sf(c, dof)


@pytest.mark.parametrize("cfg", ["NR", "ER"])
def test_band_and_spectra_at_high_statistics(ctx, ref, cfg):
    from helpers import band_chi2
    from nest_b200 import capi, shard
    cores = min(len(os.sched_getaffinity(0)), 16)
    n_ref = 300000
    with mp.get_context("spawn").Pool(cores) as pool:
        parts = pool.map(_ref_worker, [(cfg, 100 + k, n_ref) for k in range(cores)])
    ana = capi.analysis(0)
    ref_words = sum(p[0] for p in parts)
    ref_h = {c: sum(p[1][c] for p in parts) for c in COLS}
    band_ref = shard.band_from_words(ana, ref_words).finalize()
    # GPU: band from 4e7 events (device accumulator), spectra from 8e6 events (host columns)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    if cfg == "NR":
        src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    else:
        src = capi.source(capi.SRC_FLAT, capi.beta, 0.1, 20.0, 180.0)
        mo.er_model = 0
    band = capi.Band(ana)
    ctx.run(det, mo, ana, src, rc, 7, 40_000_000, columns=[], band=band)
    band_gpu = band.finalize()
    cm, cw, dof = band_chi2(band_gpu, band_ref)
    assert cm < 1.6 and cw < 1.6, "%s band chi2/dof: mean %.2f width %.2f (dof %d)" % (cfg, cm, cw, dof)
    # efficiencies per bin agree (accepted / (accepted + rejected))
    ok = (band_ref[:, 5] > 0) & np.isfinite(band_ref[:, 5]) & (ref_words[:ana.numBins * 6].reshape(-1, 6)[:, 0] > 2000)
    assert np.abs(band_gpu[ok][:, 5] - band_ref[ok][:, 5]).max() < 0.01
    cols = ctx.run(det, mo, ana, src, rc, 8, 8_000_000, columns=["photons", "electrons", "s1_5", "s2_7"])
    ed = _edges(cfg)
    gh = {"Nph": np.histogram(cols["photons"], ed["Nph"])[0], "Ne": np.histogram(cols["electrons"], ed["Ne"])[0],
          "s1c": np.histogram(np.abs(cols["s1_5"]), ed["s1c"])[0], "s2c": np.histogram(np.abs(cols["s2_7"]), ed["s2c"])[0]}
    for c in COLS:
        chi, dof, p = chi2_hist(gh[c], ref_h[c])
        assert p > 1e-4, "%s %s: chi2 %.1f / %d dof (p = %.2e)" % (cfg, c, chi, dof, p)
