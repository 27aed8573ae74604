"""Wider GPU parity: the runNESTvec flavour of the chain, detectors flattened to look-up tables,
the WIMP sampler, chunked launches and size-independent properties at larger N."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import refprobe
from helpers import COLUMN_NAMES, gpu_chain, ks_report
from nest_b200 import capi, shard

pytestmark = pytest.mark.gpu
P_MIN = 1e-4


def gpu_runvec(ctx, species, E, xyz, inField, seed, det=None):
    """runNESTvec semantics (execNEST.cpp:329-409) through nb200_run: LIST source, vec_mode"""
    det = det or capi.detector("LUX_Run03")
    rc = capi.prepare(det, inField if inField >= 0 else 180.0)
    mo = capi.model(0)                     # default_* vectors, SkewnessER = -999 (FullCalculation)
    mo.massNum, mo.atomNum = det.molarMass, 54.0
    ana = capi.analysis(0)
    E = np.ascontiguousarray(E, np.float64)
    x, y, z = [np.ascontiguousarray(xyz[:, k], np.float64) for k in range(3)]
    src = capi.source(capi.SRC_LIST, species, 0.0, 0.0, inField, fixed_pos=1, vec_mode=1)
    src.list_mem_kind = capi.MEM_HOST
    src.E, src.x, src.y, src.z = E.ctypes.data, x.ctypes.data, y.ctypes.data, z.ctypes.data
    c = ctx.run(det, mo, ana, src, rc, seed, E.size)
    out = np.zeros((E.size, 23))
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = c["energy_kev"], c["x_mm"], c["y_mm"], c["z_mm"]
    out[:, 4], out[:, 5] = c["electrons"], c["photons"]
    # NESTObservableArray::store_signals takes |.| of every entry (execNEST.cpp:284-319)
    out[:, 6], out[:, 7], out[:, 8] = np.abs(np.trunc(c["s1_0"])), np.abs(np.trunc(c["s1_1"])), np.abs(np.trunc(c["s1_8"]))
    for j, k in enumerate([2, 3, 4, 5, 6, 7]):
        out[:, 9 + j] = np.abs(c["s1_%d" % k])
    for j in range(4):
        out[:, 15 + j] = np.abs(np.trunc(c["s2_%d" % j]))
    for j, k in enumerate([4, 5, 6, 7]):
        out[:, 19 + j] = np.abs(c["s2_%d" % k])
    return out


@pytest.mark.parametrize("species,inField", [(capi.beta, 180.0), (capi.NR, 180.0), (capi.NR, -1.0)])
def test_runvec_semantics(ctx, ref, species, inField):
    n = 60000
    rng = np.random.default_rng(2)
    E = np.full(n, 12.0 if species == capi.beta else 25.0)
    xyz = np.tile(np.array([[60.0, -35.0, 250.0]]), (n, 1))
    g = gpu_runvec(ctx, species, E, xyz, inField, 5)
    m0 = capi.model(0)
    r = ref.runvec(species, E, xyz, inField, 6, list(m0.ERYieldsParam), list(m0.NRYieldsParam),
                   list(m0.NRERWidthsParam)[:13])[:, :23]
    for j in range(4, 23):
        a, b = g[:, j], r[:, j]
        if np.all(a == a[0]) and np.all(b == b[0]):
            assert a[0] == b[0]
            continue
        p = stats.ks_2samp(a, b).pvalue
        assert p > P_MIN, (j, p, a.mean(), b.mean())
    # a varied list: positions and energies pass through unchanged, one record per entry, in order
    E2 = rng.uniform(1.0, 90.0, 5000)
    xyz2 = np.column_stack([rng.uniform(-120, 120, 5000), rng.uniform(-120, 120, 5000), rng.uniform(60, 520, 5000)])
    g2 = gpu_runvec(ctx, species, E2, xyz2, inField, 7)
    assert np.array_equal(g2[:, 0], E2) and np.array_equal(g2[:, 1:4], xyz2)


def test_lut_detector_matches_closed_form(ctx, ref):
    """a detector flattened to (r,z) look-up tables by include/nest_b200_flatten.hh (compiled
    against the real VDetector) gives the same distributions as the closed-form LUX maps"""
    lut = capi.Detector()
    ref.L.ref_flatten_lux(C.byref(lut), 1)
    assert lut.FitS1.kind == capi.MAP_LUT_RZ and lut.FitS2.kind == capi.MAP_LUT_RZ
    n = 80000
    a, _ = gpu_chain(ctx, 3, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0)
    b, _ = gpu_chain(ctx, 4, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0, det=lut)
    ks = ks_report(a, b)
    bad = {k: v for k, v in ks.items() if v < P_MIN}
    assert not bad, bad
    # and the same events: table interpolation moves S1c by well under the statistical error
    c, _ = gpu_chain(ctx, 3, n, capi.beta, capi.SRC_FLAT, 0.1, 20.0, er_model=0, det=lut)
    same = a[:, refprobe.COL["Nph"]] == c[:, refprobe.COL["Nph"]]
    assert same.mean() > 0.999


def test_wimp_source(ctx, ref):
    """50 GeV WIMP spectrum (BASELINE config 4): same table on both sides, KS on the energies"""
    lib = capi.lib()
    dp = C.POINTER(C.c_double)
    base, expo = np.zeros(100), np.zeros(100)
    integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
    assert lib.nb200_wimp_prep(50.0, 5.0, 0.0, 11, base.ctypes.data_as(dp), expo.ctypes.data_as(dp), C.byref(integ),
                               C.byref(xmax), C.byref(div)) == 0
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo, ana = capi.model(1), capi.analysis(0)
    src = capi.source(capi.SRC_WIMP, capi.WIMP, 50.0, 1e-36, 180.0)
    src.wimp_base, src.wimp_exponent = base.ctypes.data_as(dp), expo.ctypes.data_as(dp)
    src.wimp_xMax, src.wimp_divisor, src.wimp_mass, src.wimp_day = xmax.value, div.value, 50.0, 0.0
    n = 100000
    cols = ctx.run(det, mo, ana, src, rc, 2, n, columns=["energy_kev", "photons", "s1_0"])
    E = cols["energy_kev"]
    assert E.min() >= 0.0 and E.max() <= xmax.value
    # piecewise-exponential table density (what the rejection sampler targets), sampled by numpy
    rng = np.random.default_rng(1)
    xs = rng.uniform(0, xmax.value, 4000000)
    idx = np.minimum((xs * div.value).astype(int), 99)
    f = base[idx] * np.exp(-expo[idx] * xs)
    keep = rng.uniform(0, f.max(), xs.size) < f
    p = stats.ks_2samp(E[E > 0], xs[keep][:200000]).pvalue
    assert p > 1e-3, p
    assert (E == 0).mean() < 0.02            # the reference's 100-try bail-out returns 0 keV (TestSpectra.cpp:491-495)


def test_chunked_launch_equals_single_and_properties_at_size(ctx_small_chunks):
    """N beyond one internal chunk (2^22 events in this context): same integers as the sum of its parts, and the
    size-independent invariants of the band accumulator"""
    ctx = ctx_small_chunks
    ana = capi.analysis(0)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    n = (1 << 22) + 312345     # the internal chunk is 28 x 148 x 1024 = 4 243 456 events on a B200 at 2^22
    full = capi.Band(ana)
    ctx.run(det, mo, ana, src, rc, 1, n, columns=[], band=full)
    parts = capi.Band(ana)
    for lo, hi in [(0, 1 << 21), (1 << 21, (1 << 22) - 7), ((1 << 22) - 7, n)]:
        ctx.run(det, mo, ana, src, rc, 1, hi - lo, first_event=lo, columns=[], band=parts)
    assert np.array_equal(full.words(), parts.words())
    tot = int(full.accepted.sum() + full.rejected.sum())
    assert 0.3 * n < tot <= n
    assert int(full.hist2d.sum()) <= int(full.accepted.sum())
    assert full.c.n_events == n
    b = full.finalize()
    ok = full.accepted > 1000
    assert np.all(np.diff(b[ok][:, 1]) > 0)                    # mean S1 rises bin by bin
    assert np.all((b[ok][:, 2] > ana.logMin) & (b[ok][:, 2] < ana.logMax))
    assert np.all((b[ok][:, 5] > 0) & (b[ok][:, 5] <= 1))


def test_full_size_config2_is_the_sum_of_its_eight_shards(ctx):
    """BASELINE config 2 at its full size (1e8 NR events): the band of one run equals, word for word,
    the sum of the eight contiguous id ranges an 8-GPU job would simulate (nest_b200/shard.py)"""
    from nest_b200 import shard
    ana = capi.analysis(0)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    n = 100_000_000
    full = capi.Band(ana)
    ctx.run(det, mo, ana, src, rc, 1, n, columns=[], band=full)
    parts = capi.Band(ana)
    for r in range(8):
        lo, hi = shard.event_range(r, 8, n)
        ctx.run(det, mo, ana, src, rc, 1, hi - lo, first_event=lo, columns=[], band=parts)
    assert np.array_equal(full.words(), parts.words())
    assert full.c.n_events == n == parts.c.n_events
    acc = int(full.accepted.sum())
    assert abs(acc / n - 0.619) < 0.005        # fraction of 0-100 keV NR events inside the S1/S2 window
    b = full.finalize()
    assert np.all(np.diff(b[:, 1]) > 0) and np.all(b[:, 4] < 2e-3)   # every bin populated: mean error < 0.002 dex


def test_packed_f32_record_is_the_f64_record_rounded(ctx):
    """NB200_SOA_F32: the same events, every double column stored as float (half the bytes over PCIe)"""
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo, ana = capi.model(1), capi.analysis(0)
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    n = (1 << 20) + 4321          # more than one host chunk
    cols = ["signalE", "driftTime", "smear_x", "z_mm", "photons", "electrons", "s1_2", "s1_7", "s2_0", "s2_4", "s2_7"]
    a = ctx.run(det, mo, ana, src, rc, 3, n, columns=cols)
    b = ctx.run(det, mo, ana, src, rc, 3, n, columns=cols, f32=True)
    for k in cols:
        if k in capi.SOA_I32:
            assert b[k].dtype == np.int32 and np.array_equal(a[k], b[k]), k
        else:
            assert b[k].dtype == np.float32 and np.array_equal(a[k].astype(np.float32), b[k]), k


def test_band_over_s2_uses_the_coarser_sum_scale(ctx):
    """useS2 = 2 bins the band in S2 (up to maxS2 = 9e4): the sum of X is kept in 2^-13 units there, and the
    accumulator still equals the host recomputation from the per-event signals, word for word"""
    from nest_b200 import shard
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    mo = capi.model(1)
    ana = capi.analysis(0)
    ana.useS2, ana.maxS2, ana.minS2, ana.numBins = 2, 9e4, 100.0, 60      # y = log10(S2 / S1), binned in S2
    src = capi.source(capi.SRC_FLAT, capi.NR, 0.0, 100.0, 180.0)
    band = capi.Band(ana)
    cols = ctx.run(det, mo, ana, src, rc, 8, 300000, columns=["signal1", "signal2"], band=band)
    words = shard.band_words_from_signals(ana, det.coinLevel, cols["signal1"], cols["signal2"])
    nb = ana.numBins
    got = band.words()
    per_g = np.stack([got[i * nb:(i + 1) * nb] for i in range(6)], axis=1)   # Band.words(): column-major
    per_h = words[:nb * 6].reshape(nb, 6)
    assert np.array_equal(per_g[:, :2], per_h[:, :2])
    assert np.array_equal(per_g[:, 2], per_h[:, 2])      # sum of X in 2^-13 units
    assert capi.lib().nb200_band_scale_x(C.byref(ana)) == 2.0 ** 13
    assert np.isfinite(band.finalize()[band.accepted > 100]).all()
