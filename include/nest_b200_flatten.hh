// nest_b200_flatten.hh -- header-only bridge from a reference VDetector subclass to
// the POD the C-ABI takes.  This is the reference-side binding of INTEGRATION.md: it
// is compiled by the *caller* against the reference's own
// include/Detectors/VDetector.hh (it is a template so that this repo never needs
// the reference headers).  It reads the 29 scalars + 3 noise arrays through the
// public getters (VDetector.hh:25-69) and tabulates the four virtual position maps
// the hot path calls (FitS1, FitS2, FitEF, FitTBA; VDetector.hh:131-148).
//
// Shipped detectors whose maps are closed forms get the closed form (exact);
// anything else gets a bilinear (r,z) table.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "nest_b200.h"

struct nb200_flatten_storage {
  std::vector<double> s1, s2, ef, tba;  // LUT backing store; must outlive the POD
};

namespace nb200_detail {

inline void lux_run03_maps(nb200_detector* d) {
  // LUX_Run03.hh:90-105 FitS1
  nb200_map m{};
  m.kind = NB200_MAP_LUX_RUN03;
  const double s1c[7] = {307.9, 0.3071, 0.0002257, 1.1525e-7, 318.84, 307.9, 0.};
  std::memcpy(m.c, s1c, sizeof s1c);
  m.c[6] = d->radmax;
  d->FitS1 = m;
  // LUX_Run03.hh:129-146 FitS2
  nb200_map s2{};
  s2.kind = NB200_MAP_LUX_RUN03;
  const double s2c[10] = {9156.3, 6.22750, 0.38126, 0.017144, 0.0002474,
                          1.6953e-6, 5.6513e-9, 7.3989e-12, 9156.3, 0.};
  std::memcpy(s2.c, s2c, sizeof s2c);
  s2.c[9] = d->radmax;
  d->FitS2 = s2;
  // LUX_Run03.hh:109-125 FitEF
  nb200_map ef{};
  ef.kind = NB200_MAP_LUX_RUN03;
  const double efc[6] = {158.92, 0.2209000, 0.0024485, 8.7098e-6, 1.5049e-8, 1.0110e-11};
  std::memcpy(ef.c, efc, sizeof efc);
  d->FitEF = ef;
  // LUX_Run03.hh:148-167 FitTBA()[1]
  nb200_map tba{};
  tba.kind = NB200_MAP_CONST;
  tba.c[0] = 0.449;
  d->FitTBA = tba;
}

template <class F>
inline void tabulate(nb200_map* m, std::vector<double>* store, int nr, int nz, double rmax,
                     double zmax, F f) {
  store->assign(size_t(nr) * size_t(nz), 0.);
  bool constant = true;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nz; ++j) {
      double r = nr > 1 ? rmax * double(i) / double(nr - 1) : 0.;
      double z = nz > 1 ? zmax * double(j) / double(nz - 1) : 0.;
      double v = f(r, z);
      (*store)[size_t(i) * nz + j] = v;
      if (v != (*store)[0]) constant = false;
    }
  std::memset(m, 0, sizeof *m);
  if (constant) {
    m->kind = NB200_MAP_CONST;
    m->c[0] = (*store)[0];
    store->clear();
    return;
  }
  m->kind = NB200_MAP_LUT_RZ;
  m->n0 = nr;
  m->n1 = nz;
  m->c[0] = 0.;
  m->c[1] = rmax;
  m->c[2] = 0.;
  m->c[3] = zmax;
  m->lut = store->data();
}

}  // namespace nb200_detail

// scalars and noise arrays only (VDetector.hh:25-126 getters); maps untouched
template <class Det>
inline void nb200_flatten_scalars(Det& det, nb200_detector* out) {
  out->g1 = det.get_g1();
  out->sPEres = det.get_sPEres();
  out->sPEthr = det.get_sPEthr();
  out->sPEeff = det.get_sPEeff();
  for (int i = 0; i < 4; ++i) out->noiseBaseline[i] = det.get_noiseBaseline()[i];
  out->P_dphe = det.get_P_dphe();
  out->coinWind = det.get_coinWind();
  out->coinLevel = det.get_coinLevel();
  out->numPMTs = det.get_numPMTs();
  out->OldW13eV = det.get_OldW13eV() ? 1 : 0;
  out->inGas = det.get_inGas() ? 1 : 0;
  for (int i = 0; i < 2; ++i) {
    out->noiseLinear[i] = det.get_noiseLinear()[i];
    out->noiseQuadratic[i] = det.get_noiseQuadratic()[i];
  }
  out->g1_gas = det.get_g1_gas();
  out->s2Fano = det.get_s2Fano();
  out->s2_thr = det.get_s2_thr();
  out->E_gas = det.get_E_gas();
  out->eLife_us = det.get_eLife_us();
  out->T_Kelvin = det.get_T_Kelvin();
  out->p_bar = det.get_p_bar();
  out->dtCntr = det.get_dtCntr();
  out->dt_min = det.get_dt_min();
  out->dt_max = det.get_dt_max();
  out->radius = det.get_radius();
  out->radmax = det.get_radmax();
  out->TopDrift = det.get_TopDrift();
  out->anode = det.get_anode();
  out->cathode = det.get_cathode();
  out->gate = det.get_gate();
  out->PosResFlat = det.get_PosResFlat();
  out->PosResBase = det.get_PosResBase();
  out->molarMass = det.get_molarMass();
}

// Det: any class with the VDetector public interface.  fold: Det::fold (the LCE enum
// value passed by GetS1/GetS2; the shipped detectors ignore it).
template <class Det>
inline void nb200_flatten(Det& det, nb200_detector* out, nb200_flatten_storage* st,
                          int nr = 257, int nz = 513, bool force_lut = false) {
  std::memset(out, 0, sizeof *out);
  nb200_flatten_scalars(det, out);

  if (!force_lut && det.getName() == std::string("LUX Run03")) {
    nb200_detail::lux_run03_maps(out);
    return;
  }
  const double rmax = out->radmax + 100.;  // xyResolution smear is capped at 100 mm
  const double zmax = out->TopDrift;
  nb200_detail::tabulate(&out->FitS1, &st->s1, nr, nz, rmax, zmax,
                         [&](double r, double z) { return det.FitS1(r, 0., z, Det::fold); });
  nb200_detail::tabulate(&out->FitS2, &st->s2, 4 * nr, 1, rmax, 0.,
                         [&](double r, double) { return det.FitS2(r, 0., Det::fold); });
  nb200_detail::tabulate(&out->FitEF, &st->ef, nr > 16 ? 17 : nr, 4 * nz, rmax, zmax,
                         [&](double r, double z) { return det.FitEF(r, 0., z); });
  nb200_detail::tabulate(&out->FitTBA, &st->tba, 33, 33, rmax, zmax,
                         [&](double r, double z) { return det.FitTBA(r, 0., z)[1]; });
}
