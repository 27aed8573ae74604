"""C-ABI surface and host-side logic without a GPU: the library loads, exports every symbol of
include/nest_b200.h, agrees on struct layouts, computes the per-run constants like the reference,
and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nest_b200 import capi, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_lux_run03.npz")


def test_every_header_symbol_is_exported(built):
    hdr = open(os.path.join(ROOT, "include", "nest_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    assert L.nb200_abi_version() == 1


def test_struct_layouts_match(built):
    L = built.lib()
    for i, t in enumerate([capi.Map, capi.Detector, capi.Model, capi.Analysis, capi.Source, capi.RunConsts,
                           capi.SoaOut, capi.BandHist]):
        assert L.nb200_sizeof(i) == C.sizeof(t)
    assert L.nb200_sizeof(99) == -1


def test_no_cuda_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.NestB200Error) as e:
        capi.Context(0)
    assert e.value.code == capi.ENODEV


def test_prepare_matches_reference_constants(built):
    g = np.load(GOLD)
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, 180.0)
    got = [rc.rho, rc.Wq_eV, rc.alpha, rc.vD] + list(rc.g2_params) + [rc.NewMean]
    assert got == list(g["consts_180"])          # bit-identical to the compiled reference
    assert det.inGas == 0 and rc.vD_middle == rc.vD
    # SURVEY 8(c) values
    assert rc.rho == 2.8886336386968878 and rc.vD == 1.504733982956229
    assert list(rc.g2_params) == [229.14932072161764, 0.54308501274270671, 22.914932072161765,
                                  12.44475617640823, 4.25]


def test_prepare_nonuniform_field_and_errors(built, orc):
    det = capi.detector("LUX_Run03")
    rc = capi.prepare(det, -1.0, 0.1)
    rc_o = capi.RunConsts()
    assert orc.L.orc_prepare(C.byref(det), C.c_double(-1.0), C.byref(rc_o)) == 0
    assert rc.vD_middle == rc_o.vD_middle            # same O(n^2)-ordered sums as the reference loop
    bad = capi.detector("LUX_Run03")
    bad.dt_min = 1e4                                  # execNEST.cpp:955-961
    with pytest.raises(capi.NestB200Error):
        capi.prepare(bad, 180.0)
    gas = capi.detector("LUX_Run03")
    gas.T_Kelvin, gas.p_bar = 250.0, 1.0              # vapour phase: outside this path
    with pytest.raises(capi.NestB200Error):
        capi.prepare(gas, 180.0)
    with pytest.raises(capi.NestB200Error):
        capi.detector("no such detector")


def test_builtin_detectors(built, ref):
    d = capi.detector("LUX_Run03")
    f = capi.Detector()
    ref.L.ref_flatten_lux(C.byref(f), 0)    # include/nest_b200_flatten.hh compiled against the real VDetector
    for name, typ in capi.Detector._fields_:
        if typ is capi.Map:
            a, b = getattr(d, name), getattr(f, name)
            assert a.kind == b.kind and list(a.c) == list(b.c), name
        elif hasattr(typ, "_length_"):
            assert list(getattr(d, name)) == list(getattr(f, name)), name
        else:
            assert getattr(d, name) == getattr(f, name), name
    v = capi.detector("VDetector")
    assert (v.g1, v.sPEres, v.numPMTs, v.TopDrift, v.PosResFlat) == (0.0760, 0.58, 89, 150., 3.0)


def test_defaults(built):
    m0, m1 = capi.model(0), capi.model(1)
    assert list(m0.NRERWidthsParam)[:13] == [0.4, 0.4, 0.04, 0.5, 0.19, 2.25, -0.001, 0.046452, 0.45, 0.205, -0.2, 0., 0.]
    assert list(m1.NRERWidthsParam)[6:11] == [-0.0005, 0.0329, 2.629, 0.2964, -0.3343]
    assert m0.SkewnessER == -999. and m1.SkewnessER == 0.
    a0, a1 = capi.analysis(0), capi.analysis(1)
    assert (a0.usePD, a0.useS2, a0.numBins, a0.minS1, a0.maxS1, a0.logMin, a0.logMax) == (2, 0, 98, 1.5, 99.5, 0.6, 3.6)
    assert (a1.usePD, a1.useS2, a1.numBins, a1.MCtruthE) == (1, 1, 118, 1)


def test_band_finalize_matches_getband(built, orc):
    """integer accumulator + nb200_band_finalize == the reference's two-pass GetBand"""
    g = np.load(GOLD)
    ana = capi.analysis(0)
    s1, s2 = g["band_sig"]
    words = shard.band_words_from_signals(ana, 2, s1, s2)
    got = shard.band_from_words(ana, words).finalize()
    exp = g["band"]
    ok = np.isfinite(exp).all(axis=1)
    assert ok.sum() > 60
    assert np.array_equal(got[:, 0], exp[:, 0])
    assert np.allclose(got[ok][:, 1:6], exp[ok][:, 1:6], rtol=2e-6, atol=1e-7)
    assert np.allclose(got[ok][:, 6], exp[ok][:, 6], rtol=1e-3, atol=1e-4)
    # other band styles go through the same words: log10(S2) vs S1, and the S2-binned style
    for use in (1, 2):
        a = capi.analysis(0)
        a.useS2 = use
        if use == 1:
            a.logMin, a.logMax = 1.0, 5.0
        w = shard.band_words_from_signals(a, 2, s1, s2)
        got = shard.band_from_words(a, w).finalize()
        exp = orc.getband(s2, s1, ana=a) if use == 2 else orc.getband(s1, s2, ana=a)
        ok = np.isfinite(exp).all(axis=1) & (w[:a.numBins * 6].reshape(-1, 6)[:, 0] > 5)
        assert ok.sum() > 10
        assert np.allclose(got[ok][:, 1:6], exp[ok][:, 1:6], rtol=5e-6, atol=1e-6), use


def test_wimp_prep_and_poisson(built):
    base = np.zeros(100)
    expo = np.zeros(100)
    integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
    dp = C.POINTER(C.c_double)
    rc = built.lib().nb200_wimp_prep(50.0, 5.0, 0.0, 3, base.ctypes.data_as(dp), expo.ctypes.data_as(dp),
                                     C.byref(integ), C.byref(xmax), C.byref(div))
    assert rc == 0
    assert div.value == 0.2 and 50.0 <= xmax.value <= 100.0
    k = int(xmax.value * div.value)
    assert np.all(base[:k] > 0) and np.all(expo[:k] > 0)      # falling piecewise-exponential table
    assert 1e2 < integ.value < 1e9
    # too fine a step is refused like the reference throws
    assert built.lib().nb200_wimp_prep(50.0, 0.5, 0.0, 3, base.ctypes.data_as(dp), expo.ctypes.data_as(dp),
                                       C.byref(integ), C.byref(xmax), C.byref(div)) == capi.EPHYSICS
    draws = np.array([built.lib().nb200_poisson(40.0, s) for s in range(2000)], float)
    assert abs(draws.mean() - 40.0) < 0.6 and abs(draws.var() - 40.0) < 5.0


def test_event_range_partitions():
    for n in (0, 1, 7, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard.event_range(k, w, n, 10) for k in range(w)]
            assert r[0][0] == 10 and r[-1][1] == 10 + n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.event_range(2, 2, 10)


def test_ziggurat_tables_close_with_equal_areas(built):
    """host side of the hit loop's normal sampler (nb_rng.cuh: build_zig_tables): 2048 layers of equal
    area under exp(-x^2/2) including the base strip + tail, closing exactly at f = 1."""
    from scipy.special import erfc
    L = built.lib()
    n = 2048
    x, f, w = np.zeros(n + 1), np.zeros(n + 1), np.zeros(n)
    k = np.zeros(n, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.nb200_debug_zig_tables(n, p(x), p(f), p(k), p(w)) == 0
    assert L.nb200_debug_zig_tables(1024, p(x), p(f), p(k), p(w)) != 0
    R = x[1]
    assert abs(R - 4.216370409511897) < 1e-12 and x[n] == 0.0 and f[n] == 1.0
    assert (np.diff(x) < 0).all() and (np.diff(f[1:]) > 0).all()
    assert np.allclose(f[1:], np.exp(-0.5 * x[1:] ** 2), rtol=0, atol=1e-15)
    V = R * np.exp(-0.5 * R * R) + np.sqrt(np.pi / 2) * erfc(R / np.sqrt(2))
    assert abs(x[0] * f[1] - V) < 1e-15 * 10                      # base strip: x_0 f(R) = V
    area = x[1:n] * (f[2:n + 1] - f[1:n])                         # layers 1..n-1
    assert np.abs(area / V - 1).max() < 1e-9
    # the rectangles cover the density; what sticks out of it is what the wedge tests reject (0.10 %)
    assert 0 < 2 * n * V / np.sqrt(2 * np.pi) - 1 < 1.1e-3
    assert (k == np.floor(2.0 ** 28 * (x[1:] / x[:-1])).astype(np.uint32)).all() and k[n - 1] == 0
    assert np.array_equal(w, x[:-1] / 2.0 ** 28)
    # share of draws that leave the layer core (what tests/test_gpu_samplers.py expects of the device)
    assert abs((1 - (k / 2.0 ** 28).mean()) - 0.0022711) < 1e-6


@pytest.mark.parametrize("prob", [0.173, 0.2, 0.01, 0.9])
def test_alias_tables_reproduce_the_binomial_pmf(built, prob):
    """host side of the fixed-probability binomial sampler (nb_rng.cuh: build_binomial_alias): the alias
    tables of Binomial(2^b, p), b = 0..12, give back the pmf to the 2^-32 resolution of their thresholds."""
    from scipy import stats
    L = built.lib()
    ne = 8204
    thr, alias = np.zeros(ne, np.uint32), np.zeros(ne, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.nb200_debug_alias_tables(C.c_double(prob), ne, p(thr), p(alias)) == 0
    assert L.nb200_debug_alias_tables(C.c_double(0.0), ne, p(thr), p(alias)) != 0
    for b in range(13):
        m = 1 << b
        K = m + 1
        off = m - 1 + b
        t, a = thr[off:off + K].astype(np.float64), alias[off:off + K].astype(np.int64)
        assert (a >= 0).all() and (a <= m).all()
        keep = (t + (t == 2.0 ** 32 - 1)) / 2.0 ** 32        # a saturated threshold means "always keep"
        pmf = keep / K
        np.add.at(pmf, a, (1 - keep) / K)
        exact = stats.binom.pmf(np.arange(K), m, prob)
        assert abs(pmf.sum() - 1) < 1e-12
        assert np.abs(pmf - exact).max() < 2.0 ** -31, (b, np.abs(pmf - exact).max())


def test_band_scale_and_overflow_guard(built):
    """the band's sum of X is a 64-bit fixed-point integer: its scale follows the analysis range (2^-20 units while X
    stays below 1024, coarser beyond), and nb200_band_finalize refuses an accumulator whose bins hold more events than
    their sums can carry instead of returning wrapped statistics"""
    capi = built
    a0 = capi.analysis(0)
    assert capi.lib().nb200_band_scale_x(C.byref(a0)) == 2.0 ** 20
    a2 = capi.analysis(1)
    a2.useS2 = 2                      # X = S2 up to maxS2 = 9e4 < 2^17
    assert capi.lib().nb200_band_scale_x(C.byref(a2)) == 2.0 ** 13
    b = capi.Band(a0)
    b.accepted[:] = 1000
    b.sumS1[:] = np.int64(1000 * 50.0 * 2 ** 20)
    b.sumY[:] = np.int64(1000 * 2.0 * 2 ** 30)
    b.sumY2[:] = np.int64(1000 * 4.01 * 2 ** 26)
    b.sumY3[:] = np.int64(1000 * 8.1 * 2 ** 22)
    out = b.finalize()
    assert np.allclose(out[:, 1], 50.0) and np.allclose(out[:, 2], 2.0)
    b.accepted[3] = np.uint64(3_000_000_000)      # x |log10| < 3.6 x 2^30 would pass 2^63
    with pytest.raises(capi.NestB200Error):
        b.finalize()
