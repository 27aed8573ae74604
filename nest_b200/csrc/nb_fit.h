// nb_fit.h -- the chi^2 fits of the band post-processing, one implementation for host and device.
//
// rootNEST fits every S1 slice of a band with a Gaussian (examples/rootNEST.cpp:1309-1359, GetBand_Gaussian) or its
// "skewband" (:1360-1507) through ROOT's TH1::Fit, and puts two curves through the NR means with TGraph::Fit
// (:855-878).  ROOT is not part of this path; what those calls minimise is restated here: weighted least squares,
// Levenberg-Marquardt with analytic Jacobians, parameter errors from (J^T W J)^-1 at the minimum (MIGRAD's parabolic
// errors for a chi^2).  The functions are __host__ __device__: the host entry points of include/nest_b200.h
// (nb200_band_fit_gauss, nb200_hist_fit*, nb200_graph_fit) and the device kernel that fits the bands of a whole sweep
// (k_band_fit, one WARP per grid point and S1 bin) run the same statements; only the sums over a slice's bins differ:
// they are written once, over the bins a member of a cooperating group owns, and completed by the group's reduction
// (FitSerial: one member, nothing to do -- the host; FitWarp: 32 lanes, bins dealt round-robin, butterfly all-reduce,
// which leaves bit-identical totals in every lane, so the lanes walk the minimiser in step).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define NB_FIT_HD __host__ __device__ inline
#else
#define NB_FIT_HD inline
#endif

namespace nb {

constexpr int kFitMaxPar = 4;

// who sums: one thread ...
struct FitSerial {
  NB_FIT_HD static int first() { return 0; }
  NB_FIT_HD static int step() { return 1; }
  NB_FIT_HD static void sum(double*, int) {}
  NB_FIT_HD static int imin(int v) { return v; }
  NB_FIT_HD static int imax(int v) { return v; }
};
#ifdef __CUDACC__
// ... or the 32 lanes of a warp (all of them must be there: full-mask shuffles)
struct FitWarp {
  __device__ static int first() { return int(threadIdx.x & 31u); }
  __device__ static int step() { return 32; }
  __device__ static void sum(double* v, int n) {
    for (int i = 0; i < n; ++i)
      for (int m = 16; m > 0; m >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], m);
  }
  __device__ static int imin(int v) { return __reduce_min_sync(0xffffffffu, v); }
  __device__ static int imax(int v) { return __reduce_max_sync(0xffffffffu, v); }
};
#endif

// Gauss-Jordan with partial pivoting on an n x n system (n <= 4); false when singular
NB_FIT_HD bool fit_solve(int n, const double M[kFitMaxPar][kFitMaxPar], const double* v, double* x) {
  double a[kFitMaxPar][kFitMaxPar + 1];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) a[i][j] = M[i][j];
    a[i][n] = v[i];
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (!(fabs(a[piv][c]) > 0.)) return false;
    for (int j = 0; j <= n; ++j) {
      const double t = a[c][j];
      a[c][j] = a[piv][j];
      a[piv][j] = t;
    }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r][c] / a[c][c];
      for (int j = c; j <= n; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < n; ++i) x[i] = a[i][n] / a[i][i];
  return true;
}

// model 0: A exp(-(x - mu)^2 / (2 s^2))                                          (ROOT's "gaus";   p = A, mu, s)
// model 1: A / (w sqrt(2 pi)) exp(-(x - xi)^2 / (2 w^2)) (1 + erf(al (x - xi) / (w sqrt 2)))
//                                                     (rootNEST.cpp:1361-1364 "skewband"; p = A, xi, w, al)
// model 2: p0 / (x + p1) + p2 x + p3                         (rootNEST.cpp:856-857, the Woods function through the NR means)
// model 3: p0 + (p1 - p0) / (1 + (x / p2)^p3)                 (rootNEST.cpp:870-871, its sigmoid fallback)
NB_FIT_HD int fit_npar(int model) { return model == 0 ? 3 : 4; }
NB_FIT_HD bool fit_domain(int model, const double* p) {
  if (model <= 1) return p[0] > 0. && p[2] > 0.;
  if (model == 3) return p[2] > 0.;
  return true;
}
NB_FIT_HD double fit_model(int model, const double* p, double x, double* g) {
  const double kPi = 3.14159265358979323846;
  if (model == 2) {
    if (g) {
      g[0] = 1. / (x + p[1]);
      g[1] = -p[0] / ((x + p[1]) * (x + p[1]));
      g[2] = x;
      g[3] = 1.;
    }
    return p[0] / (x + p[1]) + p[2] * x + p[3];
  }
  if (model == 3) {
    const double u = pow(x / p[2], p[3]), D = 1. + u;
    if (g) {
      g[0] = 1. - 1. / D;
      g[1] = 1. / D;
      g[2] = (p[1] - p[0]) / (D * D) * u * p[3] / p[2];
      g[3] = -(p[1] - p[0]) / (D * D) * u * log(x / p[2]);
    }
    return p[0] + (p[1] - p[0]) / D;
  }
  const double z = (x - p[1]) / p[2], e = exp(-0.5 * z * z);
  if (model == 0) {
    const double f = p[0] * e;
    if (g) {
      g[0] = e;
      g[1] = f * z / p[2];
      g[2] = f * z * z / p[2];
    }
    return f;
  }
  const double G = e / (p[2] * sqrt(2. * kPi)), t = p[3] * z / sqrt(2.);
  const double E = 1. + erf(t), dE = 2. / sqrt(kPi) * exp(-t * t);  // d erf / dt
  if (g) {
    g[0] = G * E;
    g[1] = p[0] * G * (z * E - dE * p[3] / sqrt(2.)) / p[2];
    g[2] = p[0] * G * ((z * z - 1.) * E - dE * p[3] * z / sqrt(2.)) / p[2];
    g[3] = p[0] * G * dE * z / sqrt(2.);
  }
  return p[0] * G * E;
}

// Point sets.  A histogram row as TH1::Fit sees it by default: the function at the bin centres, weights 1 / content,
// empty bins left out (weight 0 here: fit_chi2 skips them).  T: uint64_t (host API) or unsigned long long (the
// device accumulator); G: who sums (above).
template <class T, class G = FitSerial>
struct FitRow {
  typedef G Group;
  const T* row;
  int lb;
  double x0, w;
  NB_FIT_HD int size() const { return lb; }
  NB_FIT_HD double x(int k) const { return x0 + (double(k) + 0.5) * w; }
  NB_FIT_HD double y(int k) const { return double(row[k]); }
  NB_FIT_HD double wgt(int k) const { return row[k] ? 1. / double(row[k]) : 0.; }
  NB_FIT_HD int used() const {
    double n = 0.;
    for (int k = G::first(); k < lb; k += G::step()) n += row[k] ? 1. : 0.;
    G::sum(&n, 1);
    return int(n);
  }
};
struct FitPoints {
  typedef FitSerial Group;
  const double *xs, *ys, *ws;
  int n;
  NB_FIT_HD int size() const { return n; }
  NB_FIT_HD double x(int k) const { return xs[k]; }
  NB_FIT_HD double y(int k) const { return ys[k]; }
  NB_FIT_HD double wgt(int k) const { return ws[k]; }
  NB_FIT_HD int used() const { return n; }
};

// sum of w_k (y_k - f(x_k))^2 and, on request, J^T W J and J^T W r
template <class Pts>
NB_FIT_HD double fit_chi2(int model, const Pts& pts, const double* p, double JtJ[kFitMaxPar][kFitMaxPar], double* Jtr) {
  typedef typename Pts::Group G;
  const int n = fit_npar(model);
  // acc: chi^2, J^T W r [n], the upper triangle of J^T W J row by row
  double acc[1 + kFitMaxPar + kFitMaxPar * (kFitMaxPar + 1) / 2], g[kFitMaxPar];
  const int nacc = JtJ ? 1 + n + n * (n + 1) / 2 : 1;
  for (int i = 0; i < nacc; ++i) acc[i] = 0.;
  const int m = pts.size();
  for (int k = G::first(); k < m; k += G::step()) {
    const double w = pts.wgt(k);
    if (!(w > 0.)) continue;
    const double r = pts.y(k) - fit_model(model, p, pts.x(k), JtJ ? g : nullptr);
    acc[0] += w * r * r;
    if (JtJ) {
      int t = 1 + n;
      for (int i = 0; i < n; ++i) {
        acc[1 + i] += w * g[i] * r;
        for (int j = i; j < n; ++j) acc[t++] += w * g[i] * g[j];
      }
    }
  }
  G::sum(acc, nacc);
  if (JtJ) {
    int t = 1 + n;
    for (int i = 0; i < n; ++i) {
      Jtr[i] = acc[1 + i];
      for (int j = i; j < n; ++j) JtJ[i][j] = JtJ[j][i] = acc[t++];
    }
  }
  return acc[0];
}

// Levenberg-Marquardt on weighted points; returns the number of points used, 0 when there are too few or no minimum
// was found.  err[npar]; cov (or null): (J^T W J)^-1 row-major [npar][npar].
template <class Pts>
NB_FIT_HD int fit_lm(int model, const Pts& pts, double* p, double* err, double* chi2, double* cov) {
  const int n = fit_npar(model);
  const int npts = pts.used();
  if (npts <= n || !fit_domain(model, p)) return 0;
  double JtJ[kFitMaxPar][kFitMaxPar], Jtr[kFitMaxPar], lambda = 1e-3;
  double chi = fit_chi2(model, pts, p, JtJ, Jtr);
  if (!isfinite(chi)) return 0;
  for (int it = 0; it < 1000; ++it) {
    double M[kFitMaxPar][kFitMaxPar], step[kFitMaxPar], q[kFitMaxPar];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) M[i][j] = JtJ[i][j] * (i == j ? 1. + lambda : 1.);
    if (!fit_solve(n, M, Jtr, step)) return 0;
    for (int i = 0; i < n; ++i) q[i] = p[i] + step[i];
    double trial = fit_domain(model, q) ? fit_chi2(model, pts, q, nullptr, nullptr) : INFINITY;
    if (!isfinite(trial)) trial = INFINITY;
    if (trial <= chi) {
      bool done = chi - trial <= 1e-13 * chi;
      for (int i = 0; i < n; ++i) done = done && fabs(step[i]) <= 1e-9 * (fabs(p[i]) + (model <= 1 ? p[2] : 1e-3));
      for (int i = 0; i < n; ++i) p[i] = q[i];
      chi = fit_chi2(model, pts, p, JtJ, Jtr);
      lambda = lambda > 1e-12 ? lambda * 0.1 : lambda;
      if (done) break;
    } else {
      lambda *= 10.;
      if (lambda > 1e12) break;  // no downhill step left: p is the minimum to rounding
    }
  }
  // covariance = (J^T W J)^-1, column by column
  for (int i = 0; i < n; ++i) {
    double unit[kFitMaxPar] = {0., 0., 0., 0.}, col[kFitMaxPar];
    unit[i] = 1.;
    if (!fit_solve(n, JtJ, unit, col) || !(col[i] > 0.)) {
      if (model <= 1) return 0;
      col[i] = 0.;  // a curve with a flat direction (a straight line has no Woods pole): the minimum stands, the error does not
    }
    err[i] = sqrt(col[i]);
    if (cov)
      for (int j = 0; j < n; ++j) cov[j * n + i] = col[j];
  }
  *chi2 = chi;
  return npts;
}

// rootNEST.cpp:1587-1602: Owen's T by the reference's own midpoint rule (2500 steps), so that leakages agree digit for digit
NB_FIT_HD double owens_t_ref(double h, double a) {
  const int nIntSteps = 2500;
  const double dx = a / nIntSteps;
  double T = 0.;
  for (int k = 0; k < nIntSteps; ++k) {
    const double x = (k + 0.5) * dx;
    T += exp(-0.5 * h * h * (1 + x * x)) / (1 + x * x) * dx;
  }
  return T / (2. * 3.14159265358979323846);
}

// rootNEST.cpp:1567-1585 (EstimateSkew) and :1367-1374: start values of the skew-Gaussian fit from a slice's mean,
// standard deviation and sample skewness (mean of ((y - mean) / sd)^3).  start = {., xi, omega, alpha} (amplitude left
// to the caller)
NB_FIT_HD void skew_start(double mean, double sd, double skew, double* start) {
  const double kPi = 3.14159265358979323846;
  if (skew == 0.) skew = 1e-9;
  skew = skew / fabs(skew) * fmin(0.7, fabs(skew));
  const double s23 = pow(fabs(skew), 2. / 3.);
  const double delta0 = skew / fabs(skew) * sqrt(kPi / 2. * s23 / (s23 + pow((4. - kPi) / 2., 2. / 3.)));
  const double alpha0 = delta0 / sqrt(1. - delta0 * delta0);
  const double dE = alpha0 / sqrt(1. + alpha0 * alpha0);
  const double omega0 = sd / sqrt(1. - 2. * dE * dE / kPi);
  start[1] = mean - omega0 * dE * sqrt(2. / kPi);
  start[2] = omega0;
  start[3] = alpha0;
}

// One S1 slice of GetBand_Gaussian (examples/rootNEST.cpp:1309-1359, skewness == 0): Gaussian fit of the slice's row of
// the 2-D histogram.  o[8] = {A, mean, sigma, mean error, sigma error, chi^2 per degree of freedom, the reference's own
// goodness figure band[j][8] (:1346-1357), points used}; zeros when fewer than four log bins are filled or no minimum
// was found.  accepted: the slice's event count (those outside the axis are ROOT's underflow bin).  Every member of
// the group G calls it and gets the same o.
template <class G, class T>
NB_FIT_HD void fit_gauss_row(const T* row, int lb, double logMin, double w, double accepted, double* o) {
  const double kPi = 3.14159265358979323846;
  for (int i = 0; i < 8; ++i) o[i] = 0.;
  // start values from the histogram's own moments (the y = 0 entries are outside the axis: ROOT's underflow)
  double mo[3] = {0., 0., 0.};  // n, sum x, sum x^2
  for (int k = G::first(); k < lb; k += G::step()) {
    const double c = double(row[k]), x = logMin + (double(k) + 0.5) * w;
    mo[0] += c;
    mo[1] += c * x;
    mo[2] += c * x * x;
  }
  G::sum(mo, 3);
  const double n = mo[0];
  if (!(n > 1.)) return;
  const double mu = mo[1] / n;
  const double var = mo[2] / n - mu * mu + w * w / 12.;
  double p[4] = {n * w / sqrt(2. * kPi * var), mu, sqrt(var), 0.}, err[4] = {0., 0., 0., 0.}, chi = 0.;
  const int used = fit_lm(0, FitRow<T, G>{row, lb, logMin, w}, p, err, &chi, nullptr);
  if (!used) return;
  o[0] = p[0];
  o[1] = p[1];
  o[2] = p[2];
  o[3] = err[1];
  o[4] = err[2];
  o[5] = used > 3 ? chi / double(used - 3) : 0.;
  // the reference's own goodness figure (:1346-1357) in its own indexing: TH1F::operator[](k) is ROOT bin k
  // (k = 0 the underflow, which holds the events pushed as y = 0), the model taken at logMin + k w
  double g = 0.;
  for (int k = G::first(); k < lb; k += G::step()) {
    const double x = logMin + double(k) * w;
    const double model = p[0] * exp(-0.5 * (x - p[1]) * (x - p[1]) / (p[2] * p[2]));
    const double content = k == 0 ? accepted - n : double(row[k - 1]);
    const float fd = float(model + content);
    const double denom = fd > 1.f ? fd : 1.f;
    g += (content - model) * (content - model) / denom;
  }
  G::sum(&g, 1);
  g /= double(lb) - 3. - 1.;
  o[6] = (g != g || g < 0. || g >= 1.7976931348623157e308) ? 0. : g;
  o[7] = double(used);
}

// One S1 slice of the skew-Gaussian band (rootNEST.cpp:1360-1507) from the accumulator: the slice's row restricted to the
// log bins whose centres lie in [mean - 3 sd, mean + 3 sd) -- the window rootNEST re-bins on (:1319-1330), widened by
// 0.5 on both sides when it reaches beyond 2 (low edge) or 3 (high edge) as there -- fitted with "skewband" from
// EstimateSkew's start values.  mean, sd, skew: GetBand's moments of the slice (skew = mean of ((y - mean) / sd)^3).
// o[12] = {A, xi, omega, alpha, xi error, omega error, alpha error, cov(xi, omega), cov(xi, alpha), cov(omega, alpha),
// chi^2 per degree of freedom, points used}; zeros when the window holds too few filled bins or no minimum was found.
template <class G, class T>
NB_FIT_HD void fit_skew_row(const T* row, int lb, double logMin, double w, double mean, double sd, double skew, double* o) {
  for (int i = 0; i < 12; ++i) o[i] = 0.;
  if (!(sd > 0.) || !isfinite(mean) || !isfinite(skew)) return;
  double lo = mean - 3. * sd, hi = mean + 3. * sd;
  if (lo > 2. || hi > 3.) {
    lo -= 0.5;
    hi += 0.5;
  }
  int k0 = lb, k1 = 0;
  double n = 0.;
  for (int k = G::first(); k < lb; k += G::step()) {
    const double x = logMin + (double(k) + 0.5) * w;
    if (x >= lo && x < hi) {
      if (k < k0) k0 = k;
      if (k + 1 > k1) k1 = k + 1;
      n += double(row[k]);
    }
  }
  k0 = G::imin(k0);
  k1 = G::imax(k1);
  G::sum(&n, 1);
  if (k1 <= k0 || !(n > 1.)) return;
  double p[4], err[4] = {0., 0., 0., 0.}, cov[16], chi = 0.;
  skew_start(mean, sd, skew, p);
  p[0] = n * w;  // the function integrates to its amplitude: events in the window times the bin width
  const int used = fit_lm(1, FitRow<T, G>{row + k0, k1 - k0, logMin + double(k0) * w, w}, p, err, &chi, cov);
  if (!used) return;
  o[0] = p[0];
  o[1] = p[1];
  o[2] = p[2];
  o[3] = p[3];
  o[4] = err[1];
  o[5] = err[2];
  o[6] = err[3];
  o[7] = cov[1 * 4 + 2];
  o[8] = cov[1 * 4 + 3];
  o[9] = cov[2 * 4 + 3];
  o[10] = used > 4 ? chi / double(used - 4) : 0.;
  o[11] = double(used);
}

// rootNEST.cpp:1587-1602 by a group: the midpoint rule's 2500 terms dealt out over the members
template <class G>
NB_FIT_HD double owens_t_ref_by(double h, double a) {
  const int nIntSteps = 2500;
  const double dx = a / nIntSteps;
  double T = 0.;
  for (int k = G::first(); k < nIntSteps; k += G::step()) {
    const double x = (k + 0.5) * dx;
    T += exp(-0.5 * h * h * (1 + x * x)) / (1 + x * x) * dx;
  }
  G::sum(&T, 1);
  return T / (2. * 3.14159265358979323846);
}

}  // namespace nb
