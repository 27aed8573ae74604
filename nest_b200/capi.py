"""ctypes binding of the C-ABI in include/nest_b200.h (libnest_b200.so).

This is the Python-side stub of INTEGRATION.md: plain structs and pointers, no torch types.
The library is built in-tree by ``nest_b200.build.build()``; loading fails loudly when it is
missing -- there is no CPU fallback behind these calls.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NB200_LIB", os.path.join(HERE, "csrc", "libnest_b200.so"))

# error codes
OK, EINVAL, ENODEV, ECUDA, EPHYSICS, ENOMEM, EFIT = 0, -1, -2, -3, -4, -5, -6
# INTERACTION_TYPE (NEST.hh:133-156)
NR, WIMP, B8, DD, AmBe, Cf, ion, gammaRay, beta, CH3T, C14, Kr83m, ppSolar, atmNu = range(14)
NoneType = 17
S1_FULL, S1_PARAMETRIC, S1_HYBRID = 0, 1, 2
MAP_CONST, MAP_LUX_RUN03, MAP_LUT_RZ = 0, 1, 2
(SRC_MONO, SRC_FLAT, SRC_GAUSS, SRC_EXP, SRC_CH3T, SRC_DD, SRC_B8, SRC_LIST, SRC_WIMP,
 SRC_AMBE, SRC_C14, SRC_CF, SRC_PPSOLAR, SRC_ATMNU) = range(14)
MEM_HOST, MEM_DEVICE = 0, 1
SOA_F32_FLAG = 1  # nb200_soa_out.flags: the packed record (float columns)
FX_S1, FX_Y, FX_Y2, FX_Y3 = 2.0 ** 20, 2.0 ** 30, 2.0 ** 26, 2.0 ** 22

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)
c_i64p = C.POINTER(C.c_int64)


class Map(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32), ("_pad", C.c_int32),
                ("c", C.c_double * 16), ("lut", c_dp)]


class Detector(C.Structure):
    _fields_ = [
        ("g1", C.c_double), ("sPEres", C.c_double), ("sPEthr", C.c_double), ("sPEeff", C.c_double),
        ("noiseBaseline", C.c_double * 4), ("P_dphe", C.c_double), ("coinWind", C.c_double),
        ("coinLevel", C.c_int32), ("numPMTs", C.c_int32), ("OldW13eV", C.c_int32), ("inGas", C.c_int32),
        ("noiseLinear", C.c_double * 2), ("noiseQuadratic", C.c_double * 2),
        ("g1_gas", C.c_double), ("s2Fano", C.c_double), ("s2_thr", C.c_double), ("E_gas", C.c_double),
        ("eLife_us", C.c_double), ("T_Kelvin", C.c_double), ("p_bar", C.c_double),
        ("dtCntr", C.c_double), ("dt_min", C.c_double), ("dt_max", C.c_double),
        ("radius", C.c_double), ("radmax", C.c_double), ("TopDrift", C.c_double), ("anode", C.c_double),
        ("cathode", C.c_double), ("gate", C.c_double), ("PosResFlat", C.c_double),
        ("PosResBase", C.c_double), ("molarMass", C.c_double),
        ("FitS1", Map), ("FitS2", Map), ("FitEF", Map), ("FitTBA", Map)]


class Model(C.Structure):
    _fields_ = [("NRYieldsParam", C.c_double * 12), ("ERYieldsParam", C.c_double * 10),
                ("NRERWidthsParam", C.c_double * 17), ("n_widths", C.c_int32), ("er_model", C.c_int32),
                ("SkewnessER", C.c_double), ("multFact", C.c_double), ("massNum", C.c_double),
                ("atomNum", C.c_double)]


class Analysis(C.Structure):
    _fields_ = [("usePD", C.c_int32), ("useS2", C.c_int32), ("numBins", C.c_int32), ("logBins", C.c_int32),
                ("minS1", C.c_double), ("maxS1", C.c_double), ("minS2", C.c_double), ("maxS2", C.c_double),
                ("logMin", C.c_double), ("logMax", C.c_double), ("MCtruthE", C.c_int32),
                ("MCtruthPos", C.c_int32), ("s1mode", C.c_int32), ("s2mode", C.c_int32),
                ("z_step", C.c_double)]


class Source(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species", C.c_int32), ("eMin", C.c_double), ("eMax", C.c_double),
                ("par", C.c_double * 8), ("fixed_pos", C.c_int32), ("vec_mode", C.c_int32),
                ("xyz", C.c_double * 3), ("inField", C.c_double), ("list_mem_kind", C.c_int32),
                ("_pad", C.c_int32), ("E", C.c_void_p), ("x", C.c_void_p), ("y", C.c_void_p),
                ("z", C.c_void_p), ("wimp_base", c_dp), ("wimp_exponent", c_dp),
                ("wimp_xMax", C.c_double), ("wimp_divisor", C.c_double), ("wimp_mass", C.c_double),
                ("wimp_day", C.c_double)]


class RunConsts(C.Structure):
    _fields_ = [("rho", C.c_double), ("Wq_eV", C.c_double), ("alpha", C.c_double), ("vD", C.c_double),
                ("vD_middle", C.c_double), ("g2_params", C.c_double * 5), ("NewMean", C.c_double),
                ("field", C.c_double)]


SOA_F64 = ["energy_kev", "x_mm", "y_mm", "z_mm", "field", "driftTime", "smear_x", "smear_y"]
SOA_I32 = ["photons", "electrons", "ions", "excitons"]
SOA_TAIL = ["signal1", "signal2", "signalE", "Nph_mean", "Ne_mean", "deltaT_ns", "keVee"]


class SoaOut(C.Structure):
    _fields_ = ([("mem_kind", C.c_int32), ("flags", C.c_int32)] + [(n, C.c_void_p) for n in SOA_F64] +
                [(n, C.c_void_p) for n in SOA_I32] + [("s1", C.c_void_p * 9), ("s2", C.c_void_p * 9)] +
                [(n, C.c_void_p) for n in SOA_TAIL])


class BandHist(C.Structure):
    _fields_ = [("numBins", C.c_int32), ("logBins", C.c_int32), ("n_events", C.c_uint64),
                ("accepted", c_u64p), ("rejected", c_u64p), ("sumS1", c_i64p), ("sumY", c_i64p),
                ("sumY2", c_i64p), ("sumY3", c_i64p), ("hist2d", c_u64p)]


EXPORTS = [
    "nb200_abi_version", "nb200_sizeof", "nb200_create", "nb200_destroy", "nb200_last_error", "nb200_set_stream",
    "nb200_synchronize", "nb200_detector_builtin", "nb200_model_default", "nb200_analysis_default",
    "nb200_prepare", "nb200_wimp_prep", "nb200_wimp_prep_iso", "nb200_wimp_drate", "nb200_poisson", "nb200_yields", "nb200_quanta", "nb200_s2", "nb200_s1", "nb200_widths", "nb200_spike", "nb200_xy_resolution", "nb200_spectrum", "nb200_run", "nb200_run_sweep", "nb200_sweep_band_post", "nb200_sweep_band_fit", "nb200_band_fit_device", "nb200_run_async",
    "nb200_band_reset", "nb200_band_attach", "nb200_fp64_peak", "nb200_timing_enable", "nb200_timing_read", "nb200_se_stats", "nb200_band_fetch", "nb200_band_device_buffer", "nb200_allreduce_hist",
    "nb200_band_hist_bytes", "nb200_band_scale_x", "nb200_band_finalize", "nb200_band_chi2", "nb200_band_leakage", "nb200_band_fit_gauss", "nb200_band_leakage_gauss", "nb200_hist_fit", "nb200_hist_fit_cov", "nb200_band_leakage_skew_cov", "nb200_graph_fit", "nb200_band_leakage_skew", "nb200_lar", "nb200_launch_count",
    "nb200_debug_hitmath", "nb200_debug_binomial", "nb200_debug_binomial_list", "nb200_debug_normal", "nb200_debug_binomial_fixed",
    "nb200_debug_zig_tables", "nb200_debug_alias_tables", "nb200_debug_maps", "nb200_drift_table", "nb200_density", "nb200_drift_velocity_liquid"]

_lib = None


class NestB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nest_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def lib():
    """Load libnest_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for this path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nb200_abi_version.restype = C.c_int
    L.nb200_sizeof.argtypes = [C.c_int]
    for i, t in enumerate([Map, Detector, Model, Analysis, Source, RunConsts, SoaOut, BandHist]):
        if L.nb200_sizeof(i) != C.sizeof(t):
            raise ImportError("ABI layout mismatch for %s: library %d, binding %d" %
                              (t.__name__, L.nb200_sizeof(i), C.sizeof(t)))
    L.nb200_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.nb200_destroy.argtypes = [vp]
    L.nb200_destroy.restype = None
    L.nb200_last_error.argtypes = [vp]
    L.nb200_last_error.restype = C.c_char_p
    L.nb200_set_stream.argtypes = [vp, vp]
    L.nb200_synchronize.argtypes = [vp]
    L.nb200_detector_builtin.argtypes = [C.c_char_p, C.POINTER(Detector)]
    L.nb200_model_default.argtypes = [C.c_int, C.POINTER(Model)]
    L.nb200_analysis_default.argtypes = [C.c_int, C.POINTER(Analysis)]
    L.nb200_prepare.argtypes = [C.POINTER(Detector), C.c_double, C.c_double, C.POINTER(RunConsts)]
    L.nb200_wimp_prep.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint64, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_wimp_prep_iso.argtypes = [C.c_double, C.c_double, C.c_double, c_ip, C.c_size_t, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_wimp_drate.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int)]
    L.nb200_wimp_drate.restype = C.c_double
    L.nb200_poisson.argtypes = [C.c_double, C.c_uint64]
    L.nb200_poisson.restype = C.c_uint64
    L.nb200_yields.argtypes = [vp, C.POINTER(Detector), C.POINTER(Model), C.c_int, C.c_size_t, vp, vp,
                               C.c_double, C.c_int, vp]
    L.nb200_quanta.argtypes = [vp, C.POINTER(Detector), C.POINTER(Model), C.c_size_t, vp, C.c_double, C.c_uint64,
                               C.c_uint64, C.c_int, vp, vp]
    L.nb200_s2.argtypes = [vp, C.POINTER(Detector), C.POINTER(RunConsts), C.c_size_t, vp, vp, vp, vp, C.c_uint64,
                           C.c_uint64, C.c_int, vp]
    L.nb200_s1.argtypes = [vp, C.POINTER(Detector), C.POINTER(RunConsts), C.c_int, C.c_size_t, vp, vp, vp, C.c_uint64,
                           C.c_uint64, C.c_int, vp]
    L.nb200_widths.argtypes = [vp, C.POINTER(Detector), C.POINTER(Model), C.c_size_t, vp, vp, vp, C.c_double, C.c_int, vp]
    L.nb200_spike.argtypes = [vp, C.POINTER(Detector), C.POINTER(RunConsts), C.c_size_t, vp, vp, C.c_uint64, C.c_uint64,
                              C.c_int, vp]
    L.nb200_xy_resolution.argtypes = [vp, C.POINTER(Detector), C.c_size_t, vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, vp]
    L.nb200_spectrum.argtypes = [vp, C.POINTER(Source), C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, vp]
    run_args = [vp, C.POINTER(Detector), C.POINTER(Model), C.POINTER(Analysis), C.POINTER(Source),
                C.POINTER(RunConsts), C.c_uint64, C.c_uint64, C.c_uint64]
    L.nb200_run.argtypes = run_args + [C.POINTER(SoaOut), C.POINTER(BandHist)]
    L.nb200_run_async.argtypes = run_args + [C.POINTER(SoaOut)]
    L.nb200_run_sweep.argtypes = [vp, C.c_size_t, C.POINTER(Detector), C.POINTER(Model), C.POINTER(Analysis),
                                  C.POINTER(Source), C.POINTER(RunConsts), C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.POINTER(BandHist)]
    L.nb200_sweep_band_post.argtypes = [vp, C.c_size_t, C.POINTER(Analysis), c_dp, C.c_int, c_dp, c_dp, c_dp, c_dp]
    L.nb200_sweep_band_fit.argtypes = [vp, C.c_size_t, C.POINTER(Analysis), C.c_int, c_dp, c_dp, c_dp]
    L.nb200_band_fit_device.argtypes = [vp, C.POINTER(Analysis), C.c_int, c_dp, c_dp, c_dp]
    L.nb200_band_reset.argtypes = [vp, C.POINTER(Analysis)]
    L.nb200_band_fetch.argtypes = [vp, C.POINTER(BandHist)]
    L.nb200_band_attach.argtypes = [vp, C.POINTER(Analysis), vp]
    L.nb200_fp64_peak.argtypes = [vp, c_dp]
    L.nb200_se_stats.argtypes = [vp, C.POINTER(Detector), C.POINTER(RunConsts), C.c_int, C.c_uint64, c_dp, c_dp]
    L.nb200_timing_enable.argtypes = [vp, C.c_int]
    L.nb200_timing_read.argtypes = [vp, c_dp, c_u64p]
    L.nb200_band_device_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.nb200_allreduce_hist.argtypes = [vp, vp]
    L.nb200_band_hist_bytes.argtypes = [C.POINTER(Analysis)]
    L.nb200_band_hist_bytes.restype = C.c_size_t
    L.nb200_band_scale_x.argtypes = [C.POINTER(Analysis)]
    L.nb200_band_scale_x.restype = C.c_double
    L.nb200_band_finalize.argtypes = [C.POINTER(Analysis), C.POINTER(BandHist), c_dp]
    L.nb200_band_chi2.argtypes = [C.c_int, C.c_int, c_dp, c_dp, C.c_int, c_dp]
    L.nb200_band_leakage.argtypes = [C.POINTER(Analysis), C.POINTER(BandHist), c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_band_fit_gauss.argtypes = [C.POINTER(Analysis), C.POINTER(BandHist), c_dp]
    L.nb200_band_leakage_gauss.argtypes = [C.c_int, c_dp, c_dp, C.c_double, c_dp]
    L.nb200_hist_fit.argtypes = [C.c_int, c_u64p, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp]
    L.nb200_hist_fit_cov.argtypes = [C.c_int, c_u64p, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp]
    L.nb200_band_leakage_skew_cov.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_graph_fit.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_band_leakage_skew.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.nb200_lar.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, C.c_double, C.c_uint64, C.c_uint64, C.c_int,
                            vp, vp]
    L.nb200_debug_hitmath.argtypes = [vp, C.c_size_t, vp, vp]
    L.nb200_debug_binomial.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int64, C.c_double, vp]
    L.nb200_debug_binomial_fixed.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_double, vp]
    L.nb200_debug_binomial_list.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_size_t, vp, C.c_double, C.c_int, vp]
    L.nb200_debug_zig_tables.argtypes = [C.c_int, vp, vp, vp, vp]
    L.nb200_debug_alias_tables.argtypes = [C.c_double, C.c_int, vp, vp]
    L.nb200_debug_normal.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_size_t, vp, vp]
    L.nb200_launch_count.argtypes = [vp]
    L.nb200_launch_count.restype = C.c_uint64
    _lib = L
    return L


def _dptr(a):
    return a.ctypes.data_as(C.c_void_p)


def band_chi2(a, b, useS2=0, free_param=2):
    """rootNEST's band goodness of fit (examples/rootNEST.cpp:600-643): a simulated, b compared against"""
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = np.zeros(6)
    rc = lib().nb200_band_chi2(a.shape[0], useS2, a.ctypes.data_as(c_dp), b.ctypes.data_as(c_dp), free_param,
                               out.ctypes.data_as(c_dp))
    if rc:
        raise NestB200Error(rc, "nb200_band_chi2: binning does not match")
    return out


def hist_fit(model, counts, lo, hi, start):
    """chi^2 fit of a Gaussian (model 0: A, mu, s) or rootNEST's skew-Gaussian (model 1: A, xi, omega, alpha) to one
    histogram on (lo, hi), as TH1::Fit does by default; returns (par, err, chi2/ndf)"""
    c = np.ascontiguousarray(counts, np.uint64)
    n = 3 if model == 0 else 4
    st = np.ascontiguousarray(start, np.float64)
    if st.size != n:
        raise ValueError("model %d takes %d start values" % (model, n))
    par, err, chi = np.zeros(n), np.zeros(n), C.c_double()
    rc = lib().nb200_hist_fit(model, c.ctypes.data_as(c_u64p), c.size, lo, hi, st.ctypes.data_as(c_dp),
                              par.ctypes.data_as(c_dp), err.ctypes.data_as(c_dp), C.byref(chi))
    if rc:
        raise NestB200Error(rc, "nb200_hist_fit: no minimum" if rc == EFIT else "nb200_hist_fit")
    return par, err, chi.value


def hist_fit_cov(model, counts, lo, hi, start):
    """hist_fit with the full covariance matrix of the parameters -> (par, cov[npar][npar], chi2/ndf)"""
    c = np.ascontiguousarray(counts, np.uint64)
    n = 3 if model == 0 else 4
    st = np.ascontiguousarray(start, np.float64)
    if st.size != n:
        raise ValueError("model %d takes %d start values" % (model, n))
    par, cov, chi = np.zeros(n), np.zeros((n, n)), C.c_double()
    rc = lib().nb200_hist_fit_cov(model, c.ctypes.data_as(c_u64p), c.size, lo, hi, st.ctypes.data_as(c_dp),
                                  par.ctypes.data_as(c_dp), cov.ctypes.data_as(c_dp), C.byref(chi))
    if rc:
        raise NestB200Error(rc, "nb200_hist_fit_cov: no minimum" if rc == EFIT else "nb200_hist_fit_cov")
    return par, cov, chi.value


def band_leakage_skew_cov(skew, line, line_err):
    """rootNEST's skewness = 2 leakage with propagated error (rootNEST.cpp:774-845): skew [n][9] -> (leak, err)"""
    k = np.ascontiguousarray(skew, np.float64)
    ln = np.ascontiguousarray(line, np.float64)
    le = np.ascontiguousarray(line_err, np.float64)
    if k.ndim != 2 or k.shape[1] != 9 or ln.size != k.shape[0] or le.size != k.shape[0]:
        raise ValueError("skew must be [n][9], line and line_err [n]")
    leak, err = np.zeros(ln.size), np.zeros(ln.size)
    rc = lib().nb200_band_leakage_skew_cov(ln.size, k.ctypes.data_as(c_dp), ln.ctypes.data_as(c_dp), le.ctypes.data_as(c_dp),
                                           leak.ctypes.data_as(c_dp), err.ctypes.data_as(c_dp))
    if rc:
        raise NestB200Error(rc, "nb200_band_leakage_skew_cov")
    return leak, err


def graph_fit(model, x, y, start):
    """unweighted least squares of rootNEST's NR-centroid curves (model 2 Woods, 3 sigmoid; rootNEST.cpp:855-878);
    returns (par[4], sum of squared residuals / (n - 4))"""
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    st = np.ascontiguousarray(start, np.float64)
    if x.size != y.size or st.size != 4:
        raise ValueError("x, y of equal length and four start values")
    par, chi = np.zeros(4), C.c_double()
    rc = lib().nb200_graph_fit(model, x.size, x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp), st.ctypes.data_as(c_dp),
                               par.ctypes.data_as(c_dp), C.byref(chi))
    if rc:
        raise NestB200Error(rc, "nb200_graph_fit: no minimum" if rc == EFIT else "nb200_graph_fit")
    return par, chi.value


def band_leakage_skew(xi, omega, omega_err, alpha, line):
    """skew-normal leakage below `line` per S1 bin with the omega +- error variants (rootNEST.cpp:700-771)"""
    a = [np.ascontiguousarray(v, np.float64) for v in (xi, omega, omega_err, alpha, line)]
    n = a[0].size
    if any(v.size != n for v in a):
        raise ValueError("arrays of different lengths")
    leak, hi, lo = np.zeros(n), np.zeros(n), np.zeros(n)
    rc = lib().nb200_band_leakage_skew(n, *[v.ctypes.data_as(c_dp) for v in a], leak.ctypes.data_as(c_dp),
                                       hi.ctypes.data_as(c_dp), lo.ctypes.data_as(c_dp))
    if rc:
        raise NestB200Error(rc, "nb200_band_leakage_skew")
    return leak, hi, lo


def band_leakage_gauss(er_fit, nr_fit, num_sig_above=0.0):
    """rootNEST's Gaussian leakage columns (examples/rootNEST.cpp:700-738): rows {numSigma, leak, +err, -err}"""
    e = np.ascontiguousarray(er_fit, np.float64)
    r = np.ascontiguousarray(nr_fit, np.float64)
    if e.shape != r.shape or e.ndim != 2 or e.shape[1] != 8:
        raise ValueError("fits must be [numBins][8] as Band.fit_gauss returns them")
    out = np.zeros((e.shape[0], 4))
    rc = lib().nb200_band_leakage_gauss(e.shape[0], e.ctypes.data_as(c_dp), r.ctypes.data_as(c_dp), num_sig_above,
                                        out.ctypes.data_as(c_dp))
    if rc:
        raise NestB200Error(rc, "nb200_band_leakage_gauss")
    return out


class Band:
    """Host-side band accumulator (integers) + GetBand statistics (execNEST.cpp:1508-1621)."""

    def __init__(self, ana):
        self.ana = ana
        nb, lb = ana.numBins, ana.logBins
        self.accepted = np.zeros(nb, np.uint64)
        self.rejected = np.zeros(nb, np.uint64)
        self.sumS1 = np.zeros(nb, np.int64)
        self.sumY = np.zeros(nb, np.int64)
        self.sumY2 = np.zeros(nb, np.int64)
        self.sumY3 = np.zeros(nb, np.int64)
        self.hist2d = np.zeros(nb * max(lb, 0), np.uint64)
        self.c = BandHist(nb, lb, 0, self.accepted.ctypes.data_as(c_u64p), self.rejected.ctypes.data_as(c_u64p),
                          self.sumS1.ctypes.data_as(c_i64p), self.sumY.ctypes.data_as(c_i64p),
                          self.sumY2.ctypes.data_as(c_i64p), self.sumY3.ctypes.data_as(c_i64p),
                          self.hist2d.ctypes.data_as(c_u64p))

    def finalize(self):
        out = np.zeros((self.ana.numBins, 7))
        rc = lib().nb200_band_finalize(C.byref(self.ana), C.byref(self.c), out.ctypes.data_as(c_dp))
        if rc:
            raise NestB200Error(rc, "nb200_band_finalize")
        return out

    def leakage(self, centroid):
        """fraction of events below `centroid[numBins]` per S1 bin, Poisson error bars, totals (rootNEST.cpp:880-915)"""
        nb = self.ana.numBins
        cen = np.ascontiguousarray(centroid, np.float64)
        leak, hi, lo, tot = np.zeros(nb), np.zeros(nb), np.zeros(nb), np.zeros(3)
        dp = lambda a: a.ctypes.data_as(c_dp)
        rc = lib().nb200_band_leakage(C.byref(self.ana), C.byref(self.c), dp(cen), dp(leak), dp(hi), dp(lo), dp(tot))
        if rc:
            raise NestB200Error(rc, "nb200_band_leakage")
        return leak, hi, lo, tot

    def fit_gauss(self):
        """Gaussian fit per S1 bin (rootNEST.cpp:1309-1359): rows {A, mean, sigma, mean err, sigma err, chi2/ndf,
        the reference's band[j][8], points}"""
        out = np.zeros((self.ana.numBins, 8))
        rc = lib().nb200_band_fit_gauss(C.byref(self.ana), C.byref(self.c), out.ctypes.data_as(c_dp))
        if rc:
            raise NestB200Error(rc, "nb200_band_fit_gauss")
        return out

    def words(self):
        """all integer words, for exact comparisons"""
        return np.concatenate([self.accepted.view(np.int64), self.rejected.view(np.int64), self.sumS1,
                               self.sumY, self.sumY2, self.sumY3, self.hist2d.view(np.int64)])


def detector(name="LUX_Run03"):
    d = Detector()
    rc = lib().nb200_detector_builtin(name.encode(), C.byref(d))
    if rc:
        raise NestB200Error(rc, lib().nb200_last_error(None).decode())
    return d


def model(which=1):
    m = Model()
    lib().nb200_model_default(which, C.byref(m))
    return m


def analysis(which=0):
    a = Analysis()
    lib().nb200_analysis_default(which, C.byref(a))
    return a


def prepare(det, inField, z_step=0.0):
    rc_ = RunConsts()
    rc = lib().nb200_prepare(C.byref(det), inField, z_step, C.byref(rc_))
    if rc:
        raise NestB200Error(rc, lib().nb200_last_error(None).decode())
    return rc_


def source(kind, species, eMin=0.0, eMax=0.0, inField=180.0, fixed_pos=0, xyz=(0.0, 0.0, 0.0), par=(),
           vec_mode=0):
    s = Source()
    s.kind, s.species, s.eMin, s.eMax, s.inField = kind, species, eMin, eMax, inField
    s.fixed_pos, s.vec_mode = fixed_pos, vec_mode
    for i, v in enumerate(xyz):
        s.xyz[i] = v
    for i, v in enumerate(par):
        s.par[i] = v
    return s


class WimpTable:
    """TestSpectra::WIMP_spectrum_prep (TestSpectra.hh:43-49) from nb200_wimp_prep (isotopes from the Philox host
    stream of `seed`) or nb200_wimp_prep_iso (isotopes given, e.g. the reference's own draws)."""

    def __init__(self, mass, eStep=5.0, day=0.0, seed=1, isotopes=None):
        self.mass, self.day = float(mass), float(day)
        self.base, self.exponent = np.zeros(100), np.zeros(100)
        integ, xmax, div = C.c_double(), C.c_double(), C.c_double()
        dp = lambda a: a.ctypes.data_as(c_dp)
        if isotopes is None:
            rc = lib().nb200_wimp_prep(mass, eStep, day, seed, dp(self.base), dp(self.exponent), C.byref(integ),
                                       C.byref(xmax), C.byref(div))
        else:
            iso = np.ascontiguousarray(isotopes, np.int32)
            rc = lib().nb200_wimp_prep_iso(mass, eStep, day, iso.ctypes.data_as(c_ip), iso.size, dp(self.base),
                                           dp(self.exponent), C.byref(integ), C.byref(xmax), C.byref(div))
        if rc:
            raise NestB200Error(rc, lib().nb200_last_error(None).decode())
        self.integral, self.xMax, self.divisor = integ.value, xmax.value, div.value

    def attach(self, src):
        """point a Source at this table (keeps the arrays alive through self)"""
        src.wimp_base, src.wimp_exponent = self.base.ctypes.data_as(c_dp), self.exponent.ctypes.data_as(c_dp)
        src.wimp_xMax, src.wimp_divisor, src.wimp_mass, src.wimp_day = self.xMax, self.divisor, self.mass, self.day
        return src


SOA_COLUMNS = (SOA_F64 + SOA_I32 + ["s1_%d" % i for i in range(9)] + ["s2_%d" % i for i in range(9)] + SOA_TAIL)


class Context:
    """One context per device (nb200_create / nb200_destroy)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().nb200_create(device, C.byref(self.h))
        if rc:
            raise NestB200Error(rc, lib().nb200_last_error(None).decode())

    def close(self):
        if self.h:
            lib().nb200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise NestB200Error(rc, lib().nb200_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(lib().nb200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(lib().nb200_synchronize(self.h))

    def launch_count(self):
        return int(lib().nb200_launch_count(self.h))

    def yields(self, det, mo, species, E, field, density):
        """GetYields on the device for host arrays E, field -> dict of 6 columns."""
        E = np.ascontiguousarray(E, np.float64)
        F = np.ascontiguousarray(np.broadcast_to(field, E.shape), np.float64)
        out = np.empty((6, E.size))
        self._check(lib().nb200_yields(self.h, C.byref(det), C.byref(mo), species, E.size, _dptr(E), _dptr(F),
                                       density, MEM_HOST, _dptr(out)))
        return out

    def quanta(self, det, mo, yields, density, seed, first_event=0):
        """GetQuanta on the device for yield records [6][n] (as `yields` returns them) -> (int32 [4][n] photons,
        electrons, ions, excitons; float64 [2][n] recombProb, Variance)"""
        y = np.ascontiguousarray(yields, np.float64)
        if y.ndim != 2 or y.shape[0] != 6:
            raise ValueError("yields must be [6][n]")
        n = y.shape[1]
        qi, qd = np.zeros((4, n), np.int32), np.zeros((2, n))
        self._check(lib().nb200_quanta(self.h, C.byref(det), C.byref(mo), n, _dptr(y), density, seed, first_event,
                                       MEM_HOST, _dptr(qi), _dptr(qd)))
        return qi, qd

    def widths(self, det, mo, efield, elecFrac, numQuanta, density):
        """RecombOmegaNR / RecombOmegaER / FanoER for n records -> [3][n]"""
        f = np.ascontiguousarray(efield, np.float64)
        e = np.ascontiguousarray(np.broadcast_to(elecFrac, f.shape), np.float64)
        q = np.ascontiguousarray(np.broadcast_to(numQuanta, f.shape), np.float64)
        out = np.zeros((3, f.size))
        self._check(lib().nb200_widths(self.h, C.byref(det), C.byref(mo), f.size, _dptr(f), _dptr(e), _dptr(q), density,
                                       MEM_HOST, _dptr(out)))
        return out

    def spike(self, det, rc, pos, scint, seed, first_event=0):
        """GetSpike for n records: pos [3][n], scint [9][n] -> [2][n]"""
        p = np.ascontiguousarray(pos, np.float64)
        sc = np.ascontiguousarray(scint, np.float64)
        n = p.shape[1]
        out = np.zeros((2, n))
        self._check(lib().nb200_spike(self.h, C.byref(det), C.byref(rc), n, _dptr(p), _dptr(sc), seed, first_event,
                                      MEM_HOST, _dptr(out)))
        return out

    def xy_resolution(self, det, x, y, A_top, seed, first_event=0):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(np.broadcast_to(y, x.shape), np.float64)
        a = np.ascontiguousarray(np.broadcast_to(A_top, x.shape), np.float64)
        out = np.zeros((2, x.size))
        self._check(lib().nb200_xy_resolution(self.h, C.byref(det), x.size, _dptr(x), _dptr(y), _dptr(a), seed,
                                              first_event, MEM_HOST, _dptr(out)))
        return out

    def spectrum(self, src, n, seed, first_event=0):
        out = np.zeros(int(n))
        self._check(lib().nb200_spectrum(self.h, C.byref(src), out.size, seed, first_event, MEM_HOST, _dptr(out)))
        return out

    def s2(self, det, rc, Ne, pos, dt, field, seed, first_event=0):
        """GetS2 (Full) on the device for n records: Ne [n], pos [5][n] (truth x, y, z, smeared x, y), dt [n], field
        (scalar or [n]) -> ionization [9][n]"""
        Ne = np.ascontiguousarray(Ne, np.int32)
        n = Ne.size
        pos = np.ascontiguousarray(pos, np.float64)
        if pos.shape != (5, n):
            raise ValueError("pos must be [5][n]")
        dt = np.ascontiguousarray(np.broadcast_to(dt, (n,)), np.float64)
        F = np.ascontiguousarray(np.broadcast_to(field, (n,)), np.float64)
        out = np.zeros((9, n))
        self._check(lib().nb200_s2(self.h, C.byref(det), C.byref(rc), n, _dptr(Ne), _dptr(pos), _dptr(dt), _dptr(F), seed,
                                   first_event, MEM_HOST, _dptr(out)))
        return out

    def s1(self, det, rc, Nph, pos, energy, seed, first_event=0, mode=None):
        """GetS1 on the device for n records: Nph [n], pos [6][n] (truth x, y, z, smeared x, y, z), energy (scalar or
        [n]), mode = S1 calculation mode (default Hybrid) -> scintillation [9][n]"""
        Nph = np.ascontiguousarray(Nph, np.int32)
        n = Nph.size
        pos = np.ascontiguousarray(pos, np.float64)
        if pos.shape != (6, n):
            raise ValueError("pos must be [6][n]")
        E = np.ascontiguousarray(np.broadcast_to(energy, (n,)), np.float64)
        out = np.zeros((9, n))
        m = analysis(0).s1mode if mode is None else mode
        self._check(lib().nb200_s1(self.h, C.byref(det), C.byref(rc), m, n, _dptr(Nph), _dptr(pos), _dptr(E), seed,
                                   first_event, MEM_HOST, _dptr(out)))
        return out

    def run(self, det, mo, ana, src, rc, seed, n, first_event=0, columns=None, band=None, f32=False):
        """nb200_run with host outputs. columns: names from SOA_COLUMNS (None = all); f32: the packed record."""
        cols = {}
        soa = SoaOut()
        soa.mem_kind = MEM_HOST
        soa.flags = SOA_F32_FLAG if f32 else 0
        ft = np.float32 if f32 else np.float64
        names = SOA_COLUMNS if columns is None else columns
        for nme in names:
            if nme in SOA_I32:
                a = np.zeros(n, np.int32)
                setattr(soa, nme, a.ctypes.data)
            elif nme.startswith("s1_") or nme.startswith("s2_"):
                a = np.zeros(n, ft)
                getattr(soa, nme[:2])[int(nme[3:])] = a.ctypes.data
            else:
                a = np.zeros(n, ft)
                setattr(soa, nme, a.ctypes.data)
            cols[nme] = a
        self._check(lib().nb200_run(self.h, C.byref(det), C.byref(mo), C.byref(ana), C.byref(src), C.byref(rc),
                                    seed, first_event, n, C.byref(soa) if names else None,
                                    C.byref(band.c) if band is not None else None))
        return cols

    def run_sweep(self, dets, models, ana, srcs, rcs, seed, n, first_event=0):
        """nb200_run_sweep over len(dets) grid points -> list of Band (one per point)"""
        k = len(dets)
        if not (len(models) == len(srcs) == len(rcs) == k):
            raise ValueError("one detector, model, source and run-constants block per grid point")
        D, M, S, R = (Detector * k)(*dets), (Model * k)(*models), (Source * k)(*srcs), (RunConsts * k)(*rcs)
        bands = [Band(ana) for _ in range(k)]
        B = (BandHist * k)(*[b.c for b in bands])
        self._check(lib().nb200_run_sweep(self.h, k, D, M, C.byref(ana), S, R, seed, first_event, n, B))
        for b, c in zip(bands, B):
            b.c.n_events = c.n_events
        return bands

    def sweep_band_post(self, n_points, ana, target=None, free_param=2, centroid=None, want_stats=True):
        """statistics / goodness of fit / leakage of the last sweep's bands, computed on the device
        -> (stats [n][numBins][7] or None, gof [n][6] or None, leak [n][3] or None)"""
        nb = ana.numBins
        tgt = None if target is None else np.ascontiguousarray(target, np.float64)
        cen = None if centroid is None else np.ascontiguousarray(centroid, np.float64)
        stats = np.zeros((n_points, nb, 7)) if want_stats else None
        gof = np.zeros((n_points, 6)) if tgt is not None else None
        leak = np.zeros((n_points, 3)) if cen is not None else None
        p = lambda a: a.ctypes.data_as(c_dp) if a is not None else None
        self._check(lib().nb200_sweep_band_post(self.h, n_points, C.byref(ana), p(tgt), free_param, p(cen), p(stats),
                                                p(gof), p(leak)))
        return stats, gof, leak

    def sweep_band_fit(self, n_points, ana, model=0, line=None):
        """per-bin Gaussian (model 0: 8 columns as band_fit_gauss) or skew-Gaussian (model 1: 12 columns) fits of the
        last sweep's bands on the device -> (fit [n][numBins][cols], leak [n][numBins] below `line` or None)"""
        ln = None if line is None else np.ascontiguousarray(line, np.float64)
        fit = np.zeros((n_points, ana.numBins, 8 if model == 0 else 12))
        leak = np.zeros((n_points, ana.numBins)) if ln is not None else None
        p = lambda a: a.ctypes.data_as(c_dp) if a is not None else None
        self._check(lib().nb200_sweep_band_fit(self.h, n_points, C.byref(ana), model, p(ln), p(fit), p(leak)))
        return fit, leak

    def band_fit_device(self, ana, model=0, line=None):
        """the same for the context's own device accumulator (after run_async) -> (fit [numBins][cols], leak or None)"""
        ln = None if line is None else np.ascontiguousarray(line, np.float64)
        fit = np.zeros((ana.numBins, 8 if model == 0 else 12))
        leak = np.zeros(ana.numBins) if ln is not None else None
        p = lambda a: a.ctypes.data_as(c_dp) if a is not None else None
        self._check(lib().nb200_band_fit_device(self.h, C.byref(ana), model, p(ln), p(fit), p(leak)))
        return fit, leak

    def band_reset(self, ana):
        self._check(lib().nb200_band_reset(self.h, C.byref(ana)))

    def run_async(self, det, mo, ana, src, rc, seed, first_event, n, soa=None):
        self._check(lib().nb200_run_async(self.h, C.byref(det), C.byref(mo), C.byref(ana), C.byref(src),
                                          C.byref(rc), seed, first_event, n, C.byref(soa) if soa else None))

    def band_attach(self, ana, device_ptr):
        self._check(lib().nb200_band_attach(self.h, C.byref(ana), C.c_void_p(device_ptr)))

    def timing_enable(self, on=True):
        self._check(lib().nb200_timing_enable(self.h, 1 if on else 0))

    def timing_read(self):
        """({'k_pre': ms, 'k_hits': ms, 'k_post': ms}, launches per stage) since the last read"""
        ms = (C.c_double * 3)()
        n = (C.c_uint64 * 3)()
        self._check(lib().nb200_timing_read(self.h, ms, n))
        return dict(zip(("k_pre", "k_hits", "k_post"), list(ms))), list(n)

    def fp64_peak(self):
        v = C.c_double()
        self._check(lib().nb200_fp64_peak(self.h, C.byref(v)))
        return v.value

    def band_fetch(self, band):
        self._check(lib().nb200_band_fetch(self.h, C.byref(band.c)))

    def band_device_buffer(self):
        p = C.c_void_p()
        n = C.c_size_t()
        self._check(lib().nb200_band_device_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def debug_hitmath(self, bits3):
        b = np.ascontiguousarray(bits3, np.uint32).reshape(-1, 3)
        out = np.empty((b.shape[0], 3))
        self._check(lib().nb200_debug_hitmath(self.h, b.shape[0], _dptr(b), _dptr(out)))
        return out

    def debug_binomial(self, n, p, count, seed=1, first_event=0):
        out = np.empty(int(count), np.int64)
        self._check(lib().nb200_debug_binomial(self.h, seed, first_event, out.size, int(n), float(p), _dptr(out)))
        return out

    def debug_binomial_list(self, n, p, which_stream, seed=1, first_event=0):
        """the in-line binomial sampler, draw i with n[i] trials on the stream of event first_event + i
        (which_stream 0: the chain's S1 hit count, 1: LArNEST's fluctuations)"""
        nn = np.ascontiguousarray(n, np.int64)
        out = np.empty(nn.size, np.int64)
        self._check(lib().nb200_debug_binomial_list(self.h, seed, first_event, nn.size, _dptr(nn), float(p),
                                                    int(which_stream), _dptr(out)))
        return out

    def debug_binomial_fixed(self, n, p, count, seed=1, first_event=0):
        out = np.empty(int(count), np.int64)
        self._check(lib().nb200_debug_binomial_fixed(self.h, seed, first_event, out.size, int(n), float(p), _dptr(out)))
        return out

    def debug_normal(self, slots, seed=1, first_event=0):
        """the hit loop's normals: three per slot (Philox block) -> (z[3*slots], kind[3*slots])"""
        out = np.empty(3 * int(slots))
        kind = np.empty(3 * int(slots), np.uint8)
        self._check(lib().nb200_debug_normal(self.h, seed, first_event, int(slots), _dptr(out), _dptr(kind)))
        return out, kind

    def lar_device(self, species, n, dE, dF, density, dY, dFl, seed=0, first_event=0):
        """LArNEST on device-resident arrays (raw pointers: E[n], field[n] in; yields[9][n], fluct[4][n] out or 0)."""
        self._check(lib().nb200_lar(self.h, species, int(n), C.c_void_p(dE), C.c_void_p(dF), density, seed, first_event,
                                    MEM_DEVICE, C.c_void_p(dY) if dY else None, C.c_void_p(dFl) if dFl else None))

    def lar(self, species, E, field, density, seed=0, first_event=0, want_fluct=True):
        E = np.ascontiguousarray(E, np.float64)
        F = np.ascontiguousarray(np.broadcast_to(field, E.shape), np.float64)
        y = np.empty((9, E.size))
        f = np.empty((4, E.size)) if want_fluct else None
        self._check(lib().nb200_lar(self.h, species, E.size, _dptr(E), _dptr(F), density, seed, first_event,
                                    MEM_HOST, _dptr(y), _dptr(f) if want_fluct else None))
        return y, f
