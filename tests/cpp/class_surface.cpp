// Calls the public model methods of NESTcalc and the TestSpectra samplers exactly as code written for the reference
// does (NEST.hh:256-333, 345, 404; TestSpectra.hh:50-69) and prints the results as %.17g, one record per line, for
// tests/test_gpu_class_surface.py to compare with the compiled reference (oracle/_ref/libnestprobe.so).
#include <cstdio>
#include <iostream>

#include "LUX_Run03_b200.hh"
#include "NEST_b200.hh"

using namespace std;
using namespace NEST;

static void put(const char* tag, const YieldResult& y) {
  printf("%s %.17g %.17g %.17g %.17g %.17g %.17g\n", tag, y.PhotonYield, y.ElectronYield, y.ExcitonRatio, y.Lindhard,
         y.ElectricField, y.DeltaT_Scint);
}

int main() {
  try {
    DetectorExample_LUX_RUN03 det;
    NESTcalc n(&det);
    const double rho = n.SetDensity(det.get_T_Kelvin(), det.get_p_bar());
    const double E[5] = {0.5, 3.0, 12.0, 41.5, 250.0}, F[3] = {50., 180., 1000.};
    for (double e : E)
      for (double f : F) {
        put("gamma", n.GetYieldGamma(e, rho, f));
        put("gamma1.2", n.GetYieldGamma(e, rho, f, 1.2));
        put("erw", n.GetYieldERWeighted(e, rho, f));
        put("nr", n.GetYieldNR(e, rho, f, 131.293));
        put("ion", n.GetYieldIon(e * 20., rho, f, 4., 2.));
        put("beta", n.GetYieldBetaGR(e, rho, f));
        put("beta0.9", n.GetYieldBetaGR(e, rho, f, default_ERYieldsParam, 0.9));
      }
    // vector forms equal the scalar ones
    vector<double> ev(E, E + 5), fv(5, 180.);
    auto v = n.GetYieldNR(ev, rho, fv, 131.293);
    for (size_t i = 0; i < v.size(); ++i) put("nrvec", v[i]);
    YieldResult bad{-3., 1e9, 0.1, 1.7, 180., -999.};
    put("valid", n.YieldResultValidity(bad, 10., 13.4));
    printf("omega %.17g %.17g %.17g\n", n.RecombOmegaNR(0.3), n.RecombOmegaER(180., 0.4, 372.),
           n.FanoER(rho, 372., 180.));
    vector<double> cli = {0.4, 0.4, 0.04, 0.50, 0.19, 2.25, -0.00050, 0.03290, 2.629, 0.2964, -.3343, 0.0000, 0.0000};
    printf("omegacli %.17g %.17g %.17g\n", n.RecombOmegaNR(0.3, cli), n.RecombOmegaER(180., 0.4, 372., cli),
           n.FanoER(rho, 372., 180., cli));
    // GetSpike / xyResolution: stochastic, the test checks their distributions
    n.SetSeed(5);
    vector<double> scint = {40., 47., 41.2, 44.0, 35.1, 37.5, 38., 40.6, 38.};
    for (int i = 0; i < 2000; ++i) {
      const vector<double>& s = n.GetSpike(300, 50., -30., 200., 1.5, 1.5, scint);
      printf("spike %.17g %.17g\n", s[0], s[1]);
    }
    vector<double> xs(4000, 60.), ys(4000, -20.), at(4000, 900.);
    for (auto& r : n.xyResolution(xs, ys, at)) printf("xy %.17g %.17g\n", r[0], r[1]);
    vector<double> one = n.xyResolution(60., -20., 900.);
    printf("xy1 %.17g %.17g\n", one[0], one[1]);
    // TestSpectra: scalar calls (served from device-drawn blocks) and a vector call
    TestSpectra::SetSeed(11);
    TestSpectra spec;
    for (int i = 0; i < 3000; ++i) printf("ch3t %.17g\n", spec.CH3T_spectrum(0., 18.6));
    for (int i = 0; i < 3000; ++i) printf("dd %.17g\n", TestSpectra::DD_spectrum(0., 80., 13., 0.12, 71.2, 20., -20.5));
    for (int i = 0; i < 3000; ++i) printf("b8 %.17g\n", TestSpectra::B8_spectrum(0., 5.));
    for (int i = 0; i < 3000; ++i) printf("ambe %.17g\n", TestSpectra::AmBe_spectrum(0.1, 200.));
    spec.wimp_spectrum_prep = TestSpectra::WIMP_prep_spectrum(50., 5., 0.);
    printf("wimpprep %.17g %.17g %.17g\n", spec.wimp_spectrum_prep.integral, spec.wimp_spectrum_prep.xMax,
           spec.wimp_spectrum_prep.divisor);
    for (int i = 0; i < 3000; ++i) printf("wimp %.17g\n", TestSpectra::WIMP_spectrum(spec.wimp_spectrum_prep, 50., 0.));
    for (double e : TestSpectra::spectrum_vec(NB200_SRC_C14, 3000, 0., 156.)) printf("c14 %.17g\n", e);
    printf("drate %.17g\n", TestSpectra::WIMP_dRate(10., 50., 0.));
  } catch (const std::exception& e) {
    cerr << "class_surface: " << e.what() << endl;
    return 3;
  }
  return 0;
}
