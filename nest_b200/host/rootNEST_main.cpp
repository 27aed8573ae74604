// rootNEST without ROOT, for the uses of examples/rootNEST.cpp that need no more than a Gaussian fit:
//
//   rootNEST_b200 NESTOutput.txt                 band of Gaussians (mode 0, one input; :1166-1195)
//   rootNEST_b200 NESTOutputER.txt NESTOutputNR.txt   leakage / discrimination table (mode 0, two inputs;
//                                                :668-950 -- examples/leakage.sh runs unchanged)
//   NB200_MODE=1 rootNEST_b200 NESTOutput.txt dataBand.txt   goodness of fit against a data band (mode 1; :587-666)
//   NB200_NUMBINS=1 rootNEST_b200 NESTOutput.txt   a mono-energetic peak: S1 / S2 / energy means and resolutions (:1044-1137)
//
// with the reference's analysis.hh defaults: skewness = 1 (skew-Gaussian fits; NB200_SKEWNESS=0 for Gaussian fits, 2 for
// the detailed table with covariances and propagated errors)
// and NRbandCenter = 3 (the NR centroid is a Woods-function curve through the per-bin means, evaluated at each
// event's S1; NB200_NRBANDCENTER=1 or 2 for the other two choices).  Not included: medians (negative NRbandCenter),
// MINUIT's IMPROVE step of skewness = 2 ("M"), modes 2 and 3 (the WIMP limits).  Inputs are execNEST's 12-column stdout (src/execNEST.cpp:1386-1405), read as GetFile does
// (:955-1036); per-bin statistics, Gaussian fits and leakage come from the library's host-side band functions
// (include/nest_b200.h: nb200_band_finalize, nb200_band_fit_gauss, nb200_hist_fit, nb200_band_leakage_gauss / _skew).  Nothing here
// simulates anything: no device is needed or used.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "NEST_b200.hh"

using namespace std;

namespace {
NEST::Analysis A;
const int freeParam = 2;       // analysis.hh:65
int skewness = 1;              // analysis.hh:67
int NRbandCenter = 3;          // analysis.hh:76: 1 the bin's mean, 3 a curve through the means, 2 the curve at the bin centre
const double NUMSIGABV = -0.;  // rootNEST.cpp:39

struct BandOfFile {
  vector<vector<double>> inputs, outputs;  // per S1 bin: S1 values, y values (GetBand's two passes, :1141-1165)
  vector<double> mom;                      // [numBins][7] nb200_band_finalize
  vector<double> fit;                      // [numBins][8] nb200_band_fit_gauss; skew fits: [1] mean, [2] sigma, [3] xi error, [4] omega error
  vector<double> xi, omega, omega_err, alpha, alpha_err;  // skew fits only (band[j][9], [11], [12], [4], [7])
  vector<double> xi_err, cov_xo, cov_xa, cov_oa;          // skewness = 2 (band[j][10], [13], [14], [15])
  size_t events = 0;
};

// GetFile :955-1036: select S1 and S2 per usePD, flag what is outside (minS, maxS) as -999
bool read_signals(const char* name, vector<double>& S1, vector<double>& S2, vector<double>* E_keV = nullptr) {
  FILE* fp = fopen(name, "r");
  if (!fp) {
    cerr << "ERROR: cannot open " << name << endl;
    return false;
  }
  char line[4096];
  while (fgets(line, sizeof line, fp)) {
    double a, b, c, d, e, f, g, h, i, j, k, l, m, n;
    if (sscanf(line, "%lf\t%lf\t%lf\t%lf,%lf,%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf", &a, &b, &c, &d, &e, &f, &g, &h,
               &i, &j, &k, &l, &m, &n) != 14)
      continue;  // header, band table or summary lines
    double s1 = -999., s2 = -999.;
    if (A.usePD <= 0 && fabs(j * 1.2) > A.minS1 && (j * 1.2) < A.maxS1)
      s1 = j * 1.2;
    else if (A.usePD == 1 && fabs(j) > A.minS1 && j < A.maxS1)
      s1 = j;
    else if (A.usePD >= 2 && fabs(k) > A.minS1 && k < A.maxS1)
      s1 = k;
    if (A.usePD <= 0 && fabs(n * 1.2) > A.minS2 && (n * 1.2) < A.maxS2)
      s2 = n * 1.2;
    else if (A.usePD >= 1 && fabs(n) > A.minS2 && n < A.maxS2)
      s2 = n;
    S1.push_back(s1);
    S2.push_back(s2);
    if (E_keV) E_keV->push_back(a);
  }
  fclose(fp);
  return true;
}

// GetBand :1198-1306 into the library's integer accumulator (fixed-point scales of include/nest_b200.h) +
// GetBand_Gaussian :1309-1359
bool band_of_file(const char* name, BandOfFile& B) {
  vector<double> S1, S2;
  if (!read_signals(name, S1, S2)) return false;
  B.events = S1.size();
  const nb200_analysis ana = A.flat();
  const int nb = ana.numBins, lb = ana.logBins;
  double binWidth, border;
  if (ana.useS2 == 2) {
    binWidth = (ana.maxS2 - ana.minS2) / double(nb);
    border = ana.minS2;
  } else {
    binWidth = (ana.maxS1 - ana.minS1) / double(nb);
    border = ana.minS1;
  }
  vector<uint64_t> acc(nb, 0), rej(nb, 0), h2d(size_t(nb) * lb, 0);
  vector<int64_t> sS1(nb, 0), sY(nb, 0), sY2(nb, 0), sY3(nb, 0);
  B.inputs.assign(nb, {});
  B.outputs.assign(nb, {});
  for (size_t i = 0; i < S1.size(); ++i) {
    if (S1[i] == 0. || S2[i] == 0.) ++rej[0];
    for (int j = 0; j < nb; ++j) {
      const double s1c = border + binWidth / 2. + double(j) * binWidth;
      if (!(fabs(S1[i]) > (s1c - binWidth / 2.) && fabs(S1[i]) <= (s1c + binWidth / 2.))) continue;
      if (S1[i] >= 0. && S2[i] >= 0.) {
        double y = 0.;
        if (S1[i] && S2[i]) y = ana.useS2 == 0 ? log10(S2[i] / S1[i]) : ana.useS2 == 1 ? log10(S2[i]) : log10(S1[i] / S2[i]);
        B.inputs[j].push_back(S1[i]);
        B.outputs[j].push_back(y);
        ++acc[j];
        sS1[j] += llrint(S1[i] * nb200_band_scale_x(&ana));
        sY[j] += llrint(y * NB200_FX_Y);
        sY2[j] += llrint(y * y * NB200_FX_Y2);
        sY3[j] += llrint(y * y * y * NB200_FX_Y3);
        if (y > ana.logMin && y < ana.logMax) {  // TH1F::Fill: the rest is under- or overflow
          int k = int((y - ana.logMin) / (ana.logMax - ana.logMin) * lb);
          k = k < 0 ? 0 : k >= lb ? lb - 1 : k;
          ++h2d[size_t(j) * lb + k];
        }
      } else
        ++rej[j];
      break;
    }
  }
  nb200_band_hist H = {nb, lb, 0, acc.data(), rej.data(), sS1.data(), sY.data(), sY2.data(), sY3.data(), h2d.data()};
  B.mom.assign(size_t(nb) * 7, 0.);
  B.fit.assign(size_t(nb) * 8, 0.);
  if (nb200_band_finalize(&ana, &H, B.mom.data()) || nb200_band_fit_gauss(&ana, &H, B.fit.data())) {
    cerr << "ERROR: band statistics failed (numBins, logBins?)" << endl;
    return false;
  }
  // a slice with entries but too few filled log bins for a three-parameter fit keeps its moments (ROOT would return
  // its start values, which are these): mean, sigma, sigma / sqrt(N), sigma / sqrt(2 N)
  for (int j = 0; j < nb; ++j) {
    double* f = &B.fit[8 * size_t(j)];
    const double* m = &B.mom[7 * size_t(j)];
    if (f[7] == 0. && acc[j] >= 2 && m[3] > 0.) {
      f[1] = m[2];
      f[2] = m[3];
      f[3] = m[4];
      f[4] = m[4] / sqrt(2.);
      f[7] = -1.;  // flag: moments, not a fit
      if (A.verbosity > 0) cerr << "Too few filled bins for a fit at S1 = " << m[0] << ": keeping the moments" << endl;
    }
  }
  if (B.events < 20000 && nb > 1 && skewness != 0) {  // :1037-1042
    skewness = 0;
    cerr << "WARNING: Not enough stats (at least 10^5 events) for skew fits so doing Gaussian" << endl;
  }
  B.xi.assign(nb, 0.);
  B.omega.assign(nb, 0.);
  B.omega_err.assign(nb, 0.);
  B.alpha.assign(nb, 0.);
  B.alpha_err.assign(nb, 0.);
  B.xi_err.assign(nb, 0.);
  B.cov_xo.assign(nb, 0.);
  B.cov_xa.assign(nb, 0.);
  B.cov_oa.assign(nb, 0.);
  bool wings = false;
  for (int j = 0; skewness && j < nb; ++j) {  // :1319-1330, :1360-1507
    double* f = &B.fit[8 * size_t(j)];
    const double mean = B.mom[7 * size_t(j) + 2], sd = B.mom[7 * size_t(j) + 3];
    for (int i = 0; i < 8; ++i) f[i] = 0.;
    const vector<double>& ys = B.outputs[j];
    if (ys.size() < 2 || !(sd > 0.)) continue;
    double lo = mean - 3. * sd, hi = mean + 3. * sd;
    if (lo > 2.) wings = true;
    if (hi > 3.) wings = true;
    if (wings) {
      lo -= 0.5;
      hi += 0.5;
    }
    vector<uint64_t> h(lb, 0);
    for (double y : ys)
      if (y >= lo && y < hi) ++h[min(lb - 1, int((y - lo) / (hi - lo) * lb))];
    // EstimateSkew :1567-1585
    double skew = 0.;
    for (double y : ys) skew += pow((y - mean) / sd, 3.);
    skew /= double(ys.size());
    if (skew == 0.) skew = 1e-9;
    skew = skew / fabs(skew) * min(0.7, fabs(skew));
    const double delta0 = skew / fabs(skew) *
                          sqrt(M_PI / 2. * pow(fabs(skew), 2. / 3.) / (pow(fabs(skew), 2. / 3.) + pow((4. - M_PI) / 2., 2. / 3.)));
    const double alpha0 = delta0 / sqrt(1. - delta0 * delta0);
    const double dE = alpha0 / sqrt(1. + alpha0 * alpha0);
    const double omega0 = sd / sqrt(1. - 2. * dE * dE / M_PI);
    // the reference starts the amplitude at the number of events (:1365); the function integrates to its amplitude, so
    // the events times the bin width is where the minimum is -- start there
    const double start[4] = {double(ys.size()) * (hi - lo) / double(lb), mean - omega0 * dE * sqrt(2. / M_PI), omega0, alpha0};
    double par[4], err[4], cov[16], chi = 0.;
    if (nb200_hist_fit_cov(1, h.data(), lb, lo, hi, start, par, cov, &chi)) {
      if (A.verbosity > 0) cerr << "Re-fitting... (stats, more? and/or logBins, fewer? might help)\n";
      continue;
    }
    for (int i = 0; i < 4; ++i) err[i] = sqrt(cov[5 * i]);
    const double d = par[3] / sqrt(1. + par[3] * par[3]);
    f[0] = par[0];
    f[1] = par[1] + par[2] * d * sqrt(2. / M_PI);          // band[j][2]
    f[2] = par[2] * sqrt(1. - 2. * d * d / M_PI);          // band[j][3]
    f[3] = err[1];                                         // band[j][5] = GetParError(1)
    f[4] = err[2];                                         // band[j][6] = GetParError(2)
    f[5] = f[6] = chi;                                     // band[j][8] = chi^2 / NDF of the fit
    f[7] = 1.;
    B.xi[j] = par[1];
    B.omega[j] = par[2];
    B.omega_err[j] = err[2];
    B.alpha[j] = par[3];
    B.alpha_err[j] = err[3];
    B.xi_err[j] = err[1];
    if (skewness == 2) {  // errors of the mean and of the standard deviation through the covariance matrix, :1419-1448
      B.cov_xo[j] = cov[1 * 4 + 2];
      B.cov_xa[j] = cov[1 * 4 + 3];
      B.cov_oa[j] = cov[2 * 4 + 3];
      // (the reference's own coefficients, kept: its d mean / d xi is omega sqrt(2/pi) delta rather than 1)
      const double dmudxi = par[2] * sqrt(2. / M_PI) * d, dmudomega = sqrt(2. / M_PI) * d,
                   dmudalpha = par[2] * sqrt(2. / M_PI) * pow(1. + par[3] * par[3], -1.5);
      f[3] = sqrt(dmudxi * dmudxi * err[1] * err[1] + dmudomega * dmudomega * err[2] * err[2] +
                  dmudalpha * dmudalpha * err[3] * err[3] + 2. * dmudxi * dmudomega * B.cov_xo[j] +
                  2. * dmudomega * dmudalpha * B.cov_oa[j] + 2. * dmudxi * dmudalpha * B.cov_xa[j]);
      const double dvardomega = 2. * par[2] * (1. - 2. / M_PI * d * d),
                   dvardalpha = -4. / M_PI * par[2] * par[2] * par[3] / (1. + par[3] * par[3]) / (1. + par[3] * par[3]);
      const double var_err = sqrt(dvardomega * dvardomega * err[2] * err[2] + dvardalpha * dvardalpha * err[3] * err[3] +
                                  2. * dvardomega * dvardalpha * B.cov_oa[j]);
      f[4] = 0.5 / f[2] * var_err;
    }
    if ((chi > 10. || f[1] > 7. || f[3] > 1. || fabs(par[3]) > 3. || err[3] > 10. || f[2] > 0.5 || f[4] > 0.1 || f[1] <= 0. ||
         f[2] <= 0.) && A.verbosity > 0)
      cerr << "Re-fitting... (stats, more? and/or logBins, fewer? might help)\n";  // :1493-1504: mode 0 reports, does not retry
  }
  if (A.verbosity > 0 && skewness == 2) {  // :1168-1179
    fprintf(stdout, "Bin Center\tBand Xi\t\tXi Err\t\tBand Omega\tOmega Err\tBand Alpha\tAlpha Err\tBand Cov Xi-Om\tBand Cov "
                    "Xi-Al\tBand Cov Om-Al\tBand Mean\tMean Err\tBand Stddev\tStddev Err\n");
    for (int o = 0; o < nb; ++o) {
      const double* f = &B.fit[8 * size_t(o)];
      fprintf(stdout, "%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\n", B.mom[7 * size_t(o)], B.xi[o],
              B.xi_err[o], B.omega[o], B.omega_err[o], B.alpha[o], B.alpha_err[o], B.cov_xo[o], B.cov_xa[o], B.cov_oa[o], f[1], f[3],
              f[2], f[4]);
    }
  } else if (A.verbosity > 0) {  // :1180-1190
    fprintf(stdout, "Bin Center\tBin Actual\tGaus Mean\tMean Error\tGaus Sigma\tSig Error\tGaus Skew\tSkew Error\tX^2/DOF\n");
    for (int o = 0; o < nb; ++o) {
      const double* m = &B.mom[7 * size_t(o)];
      const double* f = &B.fit[8 * size_t(o)];
      fprintf(stdout, "%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\n", m[0], m[1], f[1], f[3], f[2], f[4], B.alpha[o], B.alpha_err[o],
              f[6]);
    }
  }
  return true;
}

// GetFile with numBins == 1 (:1044-1137): a mono-energetic peak -- Gaussian fits of the S1, S2 and energy histograms
// (2 x NUMBINS_MAX bins around the median-ish entry), means, resolutions, their errors and chi^2 / DOF
int peak_table(const char* name) {
  vector<double> S1, S2, E;
  if (!read_signals(name, S1, S2, &E)) return 1;
  const size_t numPts = E.size();
  if (numPts < 10) {
    cerr << "ERROR: too few events in " << name << endl;
    return 1;
  }
  double eMin = 1e100, eMax = -1e100;
  for (double e : E) {
    eMin = min(eMin, e);
    eMax = max(eMax, e);
  }
  const int bins = 2 * 1000;  // 2 * NUMBINS_MAX (TestSpectra.hh:32)
  double res[3][5];           // mean, resolution %, mean error, resolution error %, chi^2 / DOF
  printf("S1 Mean\t\tS1 Res [%%]\tS2 Mean\t\tS2 Res [%%]\tEc Mean\t\tEc Res[%%]\n");
  for (int j = 0; j < 3; ++j) {
    const vector<double>& holder = j == 0 ? S1 : j == 1 ? S2 : E;
    double minimum, maximum;
    if (j == 0) {
      if (holder[numPts / 2] > 1.0) {
        minimum = holder[numPts / 2] / 2.;
        maximum = holder[numPts / 2] * 2.;
      } else {
        minimum = A.minS1;
        maximum = A.maxS1;
      }
    } else if (j == 1) {
      if (holder[numPts / 2] > 20.) {
        minimum = holder[numPts / 2] / 2.;
        maximum = holder[numPts / 2] * 2.;
      } else {
        minimum = A.minS2;
        maximum = A.maxS2;
      }
    } else {
      minimum = eMin;
      maximum = eMax;
    }
    double par[3] = {0., 0., 0.}, err[3] = {0., 0., 0.}, chi = 0.;
    for (int attempt = 0; attempt < 50; ++attempt) {  // REFIT (:1086, 1096-1104)
      if (!(maximum > minimum)) maximum = minimum + 1.;
      vector<uint64_t> h(bins, 0);
      double n = 0., sx = 0., sxx = 0.;
      for (double v : holder)
        if (v >= minimum && v < maximum) {
          ++h[min(bins - 1, int((v - minimum) / (maximum - minimum) * bins))];
          n += 1.;
          sx += v;
          sxx += v * v;
        }
      if (n < 10.) break;
      const double mu = sx / n, sd = sqrt(max(sxx / n - mu * mu, 1e-12));
      const double start[3] = {n * (maximum - minimum) / bins / (sd * sqrt(2. * M_PI)), mu, sd};
      if (nb200_hist_fit(0, h.data(), bins, minimum, maximum, start, par, err, &chi)) break;
      if (par[1] <= 0.) {
        minimum -= 10.;
        continue;
      }
      if (j == 2 && (par[1] < (eMin + eMax) / 4. || par[1] > (eMin + eMax)) && attempt == 0) {
        minimum = (eMin + eMax) / 4.;
        maximum = (eMin + eMax) * 1.;
        continue;
      }
      break;
    }
    res[j][0] = par[1];
    res[j][1] = par[2] / par[1] * 100.;
    res[j][2] = err[1];
    res[j][3] = err[2] / par[1] * 1e2;
    res[j][4] = chi;
    printf("%lf\t", res[j][0]);
    printf("%lf\t", res[j][1]);
  }
  cout << endl << "+/-" << "\t\t" << "+/-" << "\t\t" << "+/-" << "\t\t" << "+/-" << "\t\t" << "+/-" << "\t\t" << "+/-";
  printf("\n%lf\t%lf\t%lf\t%lf\t%lf\t%lf\n", res[0][2], res[0][3], res[1][2], res[1][3], res[2][2], res[2][3]);
  cout << "X^2 / DOF" << "\t\t\t" << "X^2 / DOF" << "\t\t\t" << "X^2 / DOF";
  printf("\n%lf\t\t\t%lf\t\t\t%lf\tEnd Table\n", res[0][4], res[1][4], res[2][4]);
  return 0;
}

int goodness_of_fit(const char* simFile, const char* dataFile) {  // :587-666
  const int nb = A.numBins;
  vector<double> d(size_t(nb) * 6, 0.);
  FILE* fp = fopen(dataFile, "r");
  if (!fp) {
    cerr << "ERROR: cannot open " << dataFile << endl;
    return 1;
  }
  for (int i = 0; i < nb; ++i)  // centre, actual, mean, mean error, sigma, sigma error (:603-607)
    if (fscanf(fp, "%lf %lf %lf %lf %lf %lf", &d[6 * i], &d[6 * i + 1], &d[6 * i + 2], &d[6 * i + 3], &d[6 * i + 4],
               &d[6 * i + 5]) != 6)
      break;
  fclose(fp);
  BandOfFile B;
  if (!band_of_file(simFile, B)) return 1;
  const int DoF = nb - abs(freeParam);
  double chi2[4] = {0., 0., 0., 0.};
  for (int i = 0; i < nb; ++i) {
    const double* f = &B.fit[8 * size_t(i)];
    const double* x = &d[6 * size_t(i)];
    if (fabs(B.mom[7 * size_t(i)] - x[0]) > 0.05) {
      if (A.verbosity > 0)
        cerr << "Binning doesn't match for GoF calculation. Go to analysis.hh and adjust minS1, maxS1, numBins" << endl;
      return 1;
    }
    if ((A.useS2 == 0 && fabs(f[1]) < 5.) || A.useS2 != 0) {
      double error = sqrt(pow(f[3], 2.) + pow(x[3], 2.));
      chi2[0] += pow((x[2] - f[1]) / error, 2.);
      error = sqrt(pow(f[4], 2.) + pow(x[5], 2.));
      chi2[1] += pow((x[4] - f[2]) / error, 2.);
      chi2[2] += 100. * (x[2] - f[1]) / x[2];
      chi2[3] += 100. * (x[4] - f[2]) / x[4];
    } else if (A.verbosity > 0)
      cerr << "Ignoring outliers for chi^2 calculation" << endl;
  }
  chi2[0] /= double(DoF - 1);
  chi2[2] /= nb;
  chi2[1] /= double(DoF - 1);
  chi2[3] /= nb;
  cout << "The reduced CHI^2 = " << chi2[0] << " for mean, and " << chi2[1] << " for width. ";
  cout << "Arithmetic average= " << (chi2[0] + chi2[1]) / 2. << " and geo. mean " << sqrt(chi2[0] * chi2[1])
       << " (mean+width and mean*width)" << endl;
  cout << "%-errors of the form (1/N)*sum{(NEST-data)/data} for mean and width are " << -chi2[2] << " and " << -chi2[3]
       << " (averages)" << endl;
  return 0;
}

int leakage_table(const char* file1, const char* file2) {  // :668-950
  const int nb = A.numBins;
  BandOfFile B2, B;
  cout << endl << "Calculating band for first dataset" << endl;
  if (!band_of_file(file2, B2)) return 1;
  cout << endl << "Calculating band for second dataset" << endl;
  if (!band_of_file(file1, B)) return 1;
  const bool ERis2nd = B2.fit[1] > B.fit[1];  // band2[0][2] > band[0][2]
  if (ERis2nd)
    fprintf(stderr, "\n#evts\tBinCen\tBin Actual\t#StdDev's\tLeak Frac Fit\t+ errBar\t- errBar\tDiscrim[%%]\tLower ''Half''\t+ "
                    "errBar\t- errBar\tUpperHf[%%]\n");
  else
    fprintf(stderr, "\n#evts\tBinCen\tBin Actual\t#StdDev's\tLeak Frac Fit\t+ errBar\t- errBar\tDiscrim[%%]\tLeak Frac Raw\t+ "
                    "errBar\t- errBar\tDiscrim[%%]\n");
  // the Gaussian picture: ER band against the NR line (the library takes them in that order)
  const vector<double>& er = ERis2nd ? B2.fit : B.fit;
  const vector<double>& nr = ERis2nd ? B.fit : B2.fit;
  vector<double> lk(size_t(nb) * 4, 0.);
  if (nb200_band_leakage_gauss(nb, er.data(), nr.data(), NUMSIGABV, lk.data())) return 1;
  if (skewness) {  // :700-773: the leakage and its error bars from the ER band's skew-Gaussian, the number of sigmas as before
    const BandOfFile& E = ERis2nd ? B2 : B;
    vector<double> line(nb), leak(nb), hi(nb), lo(nb);
    for (int i = 0; i < nb; ++i) line[i] = nr[8 * size_t(i) + 1] + NUMSIGABV * nr[8 * size_t(i) + 2];
    if (nb200_band_leakage_skew(nb, E.xi.data(), E.omega.data(), E.omega_err.data(), E.alpha.data(), line.data(), leak.data(),
                                hi.data(), lo.data()))
      return 1;
    vector<double> perr(nb, 0.);
    if (skewness == 2) {  // :774-845: the error propagated from the fit's covariance and the NR mean's error
      vector<double> k9(size_t(nb) * 9), lerr(nb);
      for (int i = 0; i < nb; ++i) {
        const double v[9] = {E.xi[i], E.xi_err[i], E.omega[i], E.omega_err[i], E.alpha[i], E.alpha_err[i], E.cov_xo[i], E.cov_xa[i],
                             E.cov_oa[i]};
        for (int q = 0; q < 9; ++q) k9[9 * size_t(i) + q] = v[q];
        lerr[i] = nr[8 * size_t(i) + 3];
      }
      if (nb200_band_leakage_skew_cov(nb, k9.data(), line.data(), lerr.data(), leak.data(), perr.data())) return 1;
    }
    for (int i = 0; i < nb; ++i) {
      if (!(E.omega[i] > 0.)) continue;
      lk[4 * size_t(i) + 1] = leak[i];
      lk[4 * size_t(i) + 2] = skewness == 2 ? perr[i] : fabs(hi[i] - leak[i]);
      lk[4 * size_t(i) + 3] = skewness == 2 ? perr[i] : fabs(leak[i] - lo[i]);
    }
  }
  // the NR line as a smooth curve through the per-bin means (:855-878): the Woods function, else a sigmoid, else the means
  vector<double> gx, gy;
  const vector<double>& nrMom = ERis2nd ? B.mom : B2.mom;
  for (int i = 0; i < nb; ++i)
    if (nr[8 * size_t(i) + 7] != 0.) {  // bins whose fit exists (the reference would feed the graph its empty bins too)
      gx.push_back(nrMom[7 * size_t(i)]);
      gy.push_back(nr[8 * size_t(i) + 1] + NUMSIGABV * nr[8 * size_t(i) + 2]);
    }
  int curve = 0;  // 0: none (bin means), 2: Woods, 3: sigmoid
  double cp[4] = {0., 0., 0., 0.}, chi2 = 0.;
  if (NRbandCenter != 1) {
    const double woods0[4] = {10., 15., -1.5e-3, 2.}, sig0[4] = {0.5, 3., 1e2, 0.4};
    if (!nb200_graph_fit(2, int(gx.size()), gx.data(), gy.data(), woods0, cp, &chi2) && !(chi2 > 1.5 || chi2 <= 0. || std::isnan(chi2)))
      curve = 2;
    else {
      if (A.verbosity > 0)
        cerr << "WARNING: Poor fit to NR Gaussian band centroids i.e. means of log(S2) or log(S2/S1) histograms in S1 bins. "
                "Investigate please!"
             << endl;
      if (!nb200_graph_fit(3, int(gx.size()), gx.data(), gy.data(), sig0, cp, &chi2) && !(chi2 > 2. || chi2 < 0. || std::isnan(chi2)))
        curve = 3;  // (the reference then evaluates the Woods formula with these parameters, :883-886; the fitted curve is used here)
      else if (A.verbosity > 0)
        cerr << "ERROR: Even the backup plan to use sigmoid failed as well!" << endl;
    }
  }
  auto line_at = [&](double x) {
    return curve == 2 ? cp[0] / (x + cp[1]) + cp[2] * x + cp[3] : cp[0] + (cp[1] - cp[0]) / (1. + pow(x / cp[2], cp[3]));
  };
  double finalSums[3] = {0., 0., 0.};
  for (int i = 0; i < nb; ++i) {
    // counting (:880-915): the second dataset's events below the NR line -- at the event's own S1 (NRbandCenter = 3), at
    // the bin centre (2), or the NR mean of the bin (1, or when both curve fits failed)
    const double binMean = nr[8 * size_t(i) + 1] + NUMSIGABV * nr[8 * size_t(i) + 2];
    uint64_t below = 0;
    for (size_t j = 0; j < B.outputs[i].size(); ++j) {
      double centroid = binMean;
      if (curve && NRbandCenter == 3) centroid = line_at(B.inputs[i][j]);
      if (curve && NRbandCenter == 2) centroid = line_at(nrMom[7 * size_t(i)]);
      below += B.outputs[i][j] < centroid;
    }
    const double N = double(B.outputs[i].size());
    const double leakTotal = double(below) / N;
    const double poisHi = (double(below) + sqrt(double(below))) / (N - sqrt(N));
    const double poisLo = (double(below) - sqrt(double(below))) / (N + sqrt(N));
    const double* g = &lk[4 * size_t(i)];
    cerr << B.outputs[i].size();
    fprintf(stderr, "\t%.2f\t%f\t%f\t%e\t%e\t%e\t%f\t%e\t%e\t%e\t%f\t\n", 0.5 * (B.mom[7 * size_t(i)] + B2.mom[7 * size_t(i)]),
            0.5 * (B.mom[7 * size_t(i) + 1] + B2.mom[7 * size_t(i) + 1]), g[0], g[1], g[2], g[3], (1. - g[1]) * 100., leakTotal,
            poisHi - leakTotal, leakTotal - poisLo, (1. - leakTotal) * 100.);
    finalSums[0] += double(below);
    finalSums[1] += N;
    finalSums[2] += g[1];
  }
  fprintf(stderr,
          "\nOVERALL DISCRIMINATION (ER) or NON-ACCEPTANCE (NR) between min and maxS1 = %.12f%%, total: Gaussian & "
          "non-Gaussian (tot=counting) Leakage Fraction = %.12e\n",
          (1. - finalSums[0] / finalSums[1]) * 100., finalSums[0] / finalSums[1]);
  const double LowValue = (finalSums[0] + sqrt(finalSums[0])) / (finalSums[1] - sqrt(finalSums[1]));
  const double HighValue = (finalSums[0] - sqrt(finalSums[0])) / (finalSums[1] + sqrt(finalSums[1]));
  fprintf(stderr, "highest possible discrimination value (plus  1-sigma) is %.12f%%, with corresponding leakage of %.12e\n",
          (1. - HighValue) * 100., HighValue);
  fprintf(stderr, " lowest possible discrimination value (minus 1-sigma) is %.12f%%, with corresponding leakage of %.12e\n",
          (1. - LowValue) * 100., LowValue);
  fprintf(stderr,
          "OVERALL DISCRIMINATION                             between min and maxS1 = %.12f%%, Gauss, or skew-normal fits "
          "(whatever you ran) Leakage Fraction = %.12e\n",
          (1. - finalSums[2] / nb) * 100., finalSums[2] / nb);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (const char* a = getenv("NB200_ANALYSIS"))
    if (!strcmp(a, "G2")) A = NEST::Analysis::G2();
  if (const char* v = getenv("NB200_VERBOSITY")) A.verbosity = atoi(v);
  if (const char* v = getenv("NB200_LOGBINS")) A.logBins = atoi(v);
  if (const char* v = getenv("NB200_SKEWNESS")) skewness = max(0, min(2, atoi(v)));
  if (const char* v = getenv("NB200_NRBANDCENTER")) NRbandCenter = abs(atoi(v));
  if (NRbandCenter != 1 && NRbandCenter != 2 && NRbandCenter != 3) NRbandCenter = 3;  // :101-103
  const int mode = getenv("NB200_MODE") ? atoi(getenv("NB200_MODE")) : 0;
  if (argc < 2) {  // :104-119
    cout << endl << "This program takes 1 (or 2) inputs." << endl << endl;
    cout << "One input means you are just doing a band of Gaussians." << endl << endl;
    cout << "Two inputs means you're doing both ER and NR and calculating the leakage and discrimination." << endl;
    cout << "If you write NR first, you're seeking characterization of the non-Gaussian asymmetry of the NR band." << endl;
    cout << "If you write ER first, you're finding the non-Gaussian leakage of ER into the NR band." << endl << endl;
    return 1;
  }
  if (const char* v = getenv("NB200_NUMBINS")) A.numBins = atoi(v);  // analysis.hh's numBins: 1 = a mono-energetic peak
  if (mode == 1) {
    if (argc < 3) {
      if (A.verbosity > 0) cerr << "ERROR: mode 1 requires *2* input files. 1 is not enough" << endl;
      return 1;
    }
    if (A.numBins == 1) return peak_table(argv[1]);  // :595-598
    return goodness_of_fit(argv[1], argv[2]);
  }
  if (mode != 0) {
    cerr << "ERROR: rootNEST mode " << mode << " (WIMP limits) is not part of nest-b200; use the reference rootNEST." << endl;
    return 1;
  }
  if (A.numBins == 1) return peak_table(argv[1]);
  if (argc == 2) {
    BandOfFile B;
    return band_of_file(argv[1], B) ? 0 : 1;
  }
  return leakage_table(argv[1], argv[2]);
}
