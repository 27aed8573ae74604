"""ctypes access to the UNMODIFIED reference built by oracle/Makefile into
oracle/_ref/libnestprobe.so (extern "C" probes of oracle/ref_probe.cpp over the reference's
own objects) and to the CPU restatement oracle/liboracle.so.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

from nest_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnestprobe.so")
ORC_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_BIN_LUX = os.path.join(ROOT, "oracle", "_ref", "execNEST_ref_lux")
# ... with analysis.hh's switches taken from NEST_REF_* environment variables (oracle/shim_lux_env)
REF_BIN_LUX_ENV = os.path.join(ROOT, "oracle", "_ref", "execNEST_ref_lux_env")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "execNEST_ref")  # the reference exactly as shipped (LZ_WS2024)

dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(v, n=None):
    a = np.ascontiguousarray(v, np.float64)
    return a


class Probe:
    """Common surface of the reference probe (prefix ref_) and the oracle port (prefix orc_)."""

    def __init__(self, path, prefix, needs_det):
        self.L = C.CDLL(path)
        self.px = prefix
        self.needs_det = needs_det  # oracle functions take the flattened detector first
        self.det = capi.detector("LUX_Run03")
        capi.prepare(self.det, 180.0)  # sets inGas
        self.L[prefix + "chain_ncol"].restype = C.c_int
        self.ncol = self.L[prefix + "chain_ncol"]()

    def _f(self, name):
        return self.L[self.px + name]

    def _d(self):
        return [C.byref(self.det)] if self.needs_det else []

    def yields(self, which, species, E, field, rho, massNum, atomNum, nrpar, erpar, multFact=1.0):
        E = _arr(E)
        F = _arr(np.broadcast_to(field, E.shape))
        out = np.zeros((E.size, 6))
        nr, er = _arr(nrpar), _arr(erpar)
        f = self._f("yields")
        f.restype = None if not self.needs_det else C.c_int
        f(*self._d(), C.c_int(which), C.c_int(species), C.c_int(E.size), _p(E), _p(F), C.c_double(rho),
          C.c_double(massNum), C.c_double(atomNum), _p(nr), _p(er), C.c_double(multFact), _p(out))
        return out

    def maps(self, x, y, z):
        x, y, z = _arr(x), _arr(y), _arr(z)
        o = [np.zeros(x.size) for _ in range(4)]
        self._f("maps")(*self._d(), C.c_int(x.size), _p(x), _p(y), _p(z), *[_p(a) for a in o])
        return o

    def quanta(self, seed, n, y6, rho, widths, skewER):
        """n GetQuanta draws on one fixed YieldResult -> (int [n][4] Nph, Ne, Ni, Nex; float [n][2] recombProb, Variance)"""
        y, w = _arr(y6), _arr(widths)
        io, do = np.zeros((n, 4), np.int32), np.zeros((n, 2))
        self._f("quanta")(*self._d(), C.c_uint64(seed), C.c_int(n), _p(y), C.c_double(rho), _p(w), C.c_int(w.size),
                          C.c_double(skewER), _p(io), _p(do))
        return io, do

    def s1(self, seed, n, Nph, pos, spos, vD, vDmid, species, field, energy, mode):
        """n GetS1 draws on fixed photons / positions -> scintillation [n][9]"""
        p, sp = _arr(pos), _arr(spos)
        out = np.zeros((n, 9))
        self._f("s1")(*self._d(), C.c_uint64(seed), C.c_int(n), C.c_int(Nph), _p(p), _p(sp), C.c_double(vD),
                      C.c_double(vDmid), C.c_int(species), C.c_double(field), C.c_double(energy), C.c_int(mode), _p(out))
        return out

    def s2(self, seed, n, Ne, pos, spos, dt, vD, field):
        """n GetS2 (Full) draws on fixed electrons / positions / drift time -> ionization [n][9]"""
        p, sp = _arr(pos), _arr(spos)
        out = np.zeros((n, 9))
        self._f("s2")(*self._d(), C.c_uint64(seed), C.c_int(n), C.c_int(Ne), _p(p), _p(sp), C.c_double(dt), C.c_double(vD),
                      C.c_double(field), _p(out))
        return out

    def sampler(self, kind, seed, n, p0=0.0, p1=0.0, p2=0.0):
        out = np.zeros(n)
        self._f("sampler")(C.c_int(kind), C.c_uint64(seed), C.c_int(n), C.c_double(p0), C.c_double(p1),
                           C.c_double(p2), _p(out))
        return out

    def spectrum(self, kind, seed, n, eMin, eMax):
        out = np.zeros(n)
        self._f("spectrum")(C.c_int(kind), C.c_uint64(seed), C.c_int(n), C.c_double(eMin), C.c_double(eMax),
                            _p(out))
        return out

    def chain(self, seed, n, species, src_kind, eMin, eMax, inField, fixed_pos, xyz, er_model, multFact,
              nrpar, erpar, widths, skewER, g2_verbosity=0, ana=None):
        out = np.zeros((n, self.ncol))
        xyz_, nr, er, w = _arr(xyz), _arr(nrpar), _arr(erpar), _arr(widths)
        f = self._f("chain")
        f.restype = C.c_int
        pre = self._d() + ([C.byref(ana)] if self.needs_det else [])
        rc = f(*pre, C.c_uint64(seed), C.c_int(n), C.c_int(species), C.c_int(src_kind), C.c_double(eMin),
               C.c_double(eMax), C.c_double(inField), C.c_int(fixed_pos), _p(xyz_), C.c_int(er_model),
               C.c_double(multFact), _p(nr), _p(er), _p(w), C.c_int(w.size), C.c_double(skewER),
               C.c_int(g2_verbosity), _p(out))
        if rc:
            raise RuntimeError("%schain returned %d" % (self.px, rc))
        return out

    def getband(self, s1, s2, ana=None, coinLevel=2):
        s1, s2 = _arr(s1), _arr(s2)
        out = np.zeros((1000, 7))
        f = self._f("getband")
        f.restype = C.c_int
        if self.needs_det:
            nb = f(C.byref(ana), C.c_int(coinLevel), C.c_int(s1.size), _p(s1), _p(s2), _p(out))
        else:
            nb = f(C.c_int(s1.size), _p(s1), _p(s2), _p(out))
        return out[:nb]

    def lux_set(self, g1=-1.0, g1_gas=-1.0, E_gas=-1.0):
        """reference probe only: g1 / g1_gas / E_gas of its LUX_Run03 instance (negative = header default)"""
        self.L.ref_lux_set(C.c_double(g1), C.c_double(g1_gas), C.c_double(E_gas))

    # ---- WIMP (reference probe only: TestSpectra.cpp:251-499) ----
    def wimp_drate(self, seed, ER, mass, day=0.0):
        """WIMP_dRate at ER[i] on the generator seeded seed + i -> (rate, isotope it drew)"""
        ER = _arr(ER)
        out, A = np.zeros(ER.size), np.zeros(ER.size)
        self._f("wimp_drate")(C.c_uint64(seed), C.c_int(ER.size), _p(ER), C.c_double(mass), C.c_double(day), _p(out), _p(A))
        return out, A.astype(int)

    def wimp_prep(self, seed, mass, eStep=5.0, day=0.0):
        """WIMP_prep_spectrum under `seed` -> (base[100], exponent[100], integral, xMax, divisor)"""
        base, expo, o3 = np.zeros(100), np.zeros(100), np.zeros(3)
        f = self._f("wimp_prep")
        f.restype = C.c_int
        rc = f(C.c_uint64(seed), C.c_double(mass), C.c_double(eStep), C.c_double(day), _p(base), _p(expo), _p(o3))
        if rc:
            raise RuntimeError("reference WIMP_prep_spectrum threw")
        return base, expo, o3[0], o3[1], o3[2]

    def wimp_spectrum(self, seed_prep, seed_draw, n, mass, eStep=5.0, day=0.0):
        out = np.zeros(n)
        f = self._f("wimp_spectrum")
        f.restype = C.c_int
        rc = f(C.c_uint64(seed_prep), C.c_uint64(seed_draw), C.c_int(n), C.c_double(mass), C.c_double(eStep),
               C.c_double(day), _p(out))
        if rc:
            raise RuntimeError("reference WIMP_spectrum threw")
        return out

    def poisson(self, seed, n, mean):
        out = np.zeros(n)
        self._f("poisson")(C.c_uint64(seed), C.c_int(n), C.c_double(mean), _p(out))
        return out

    def lar(self, seed, species, E, field, density):
        E = _arr(E)
        F = _arr(np.broadcast_to(field, E.shape))
        y = np.zeros((E.size, 9))
        fl = np.zeros((E.size, 4))
        self._f("lar")(C.c_uint64(seed), C.c_int(species), C.c_int(E.size), _p(E), _p(F), C.c_double(density),
                       _p(y), _p(fl))
        return y, fl

    def runvec(self, species, E, xyz, inField, seed, erpar, nrpar, widths, s1mode=2):
        E, xyz = _arr(E), _arr(xyz)
        out = np.zeros((E.size, 26))
        er, nr, w = _arr(erpar), _arr(nrpar), _arr(widths)
        f = self._f("runvec")
        f.restype = C.c_int
        f(*self._d(), C.c_int(species), C.c_int(E.size), _p(E), _p(xyz), C.c_double(inField), C.c_int(seed),
          _p(er), _p(nr), _p(w), C.c_int(w.size), C.c_int(s1mode), _p(out))
        return out


# column indices of chain output (ref_probe.cpp / nest_oracle.h)
COLS = ["keV_true", "keV_out", "field", "dt", "x", "y", "z", "sx", "sy", "Nph", "Ne", "Ni", "Nex", "Yph", "Yne"]
COL = {n: i for i, n in enumerate(COLS)}
S1_0 = len(COLS)
S2_0 = S1_0 + 9
SIG1 = S2_0 + 9
SIG2 = SIG1 + 1
SIGE = SIG2 + 1

_ref = None
_orc = None


def reference():
    """The compiled reference, or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = Probe(REF_SO, "ref_", False)
        _ref.L.ref_set_verbosity(0)
    return _ref


def oracle():
    global _orc
    if _orc is None and os.path.exists(ORC_SO):
        _orc = Probe(ORC_SO, "orc_", True)
    return _orc
