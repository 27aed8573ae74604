// A caller of the per-event stage API written against the reference's signatures (the shape of the reference's
// examples/bareNEST.cpp:183-212: yields -> quanta -> S1 -> S2 for one event), compiled against the host mirror to
// show that such code needs no edits.  Prints one line of numbers on a B200; without a CUDA device the library
// refuses (there is no CPU fallback) and this program reports that and returns 3.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "LUX_Run03_b200.hh"

using namespace NEST;
using namespace std;

int main() {
  auto* detector = new DetectorExample_LUX_RUN03();
  int status = 0;
  try {
    NESTcalc n(detector);
    const INTERACTION_TYPE type_num = NEST::beta;
    const double keV = 5., field = 180.;
    const double rho = n.SetDensity(detector->get_T_Kelvin(), detector->get_p_bar());
    const double vD = n.SetDriftVelocity(detector->get_T_Kelvin(), rho, field);
    vector<double> vTable = n.SetDriftVelocity_NonUniform(rho, 0.5, detector->get_T_Kelvin(), detector->get_p_bar(), 10., -20.);
    if (vTable.empty() || !(vTable[0] > 0.)) status = 2;
    vector<double> g2_params = n.CalculateG2(0);
    const double pos_x = 10., pos_y = -20., pos_z = 250.;
    const double driftTime = (detector->get_TopDrift() - pos_z) / vD;
    YieldResult yields = n.GetYields(type_num, keV, rho, field, 131.293, 54., default_NRYieldsParam, default_ERYieldsParam);
    QuantaResult quanta = n.GetQuanta(yields, rho, default_NRERWidthsParam, -999.);
    vector<double> wf_amp;
    vector<int64_t> wf_time;
    double truthPos[3] = {pos_x, pos_y, pos_z}, smearPos[3] = {pos_x, pos_y, pos_z};
    vector<double> scint = n.GetS1(quanta, truthPos[0], truthPos[1], truthPos[2], smearPos[0], smearPos[1], smearPos[2], vD, vD,
                                   type_num, 0, field, keV, S1CalculationMode::Hybrid, 0, wf_time, wf_amp);
    if (truthPos[2] < detector->get_cathode()) quanta.electrons = 0;
    vector<double> scint2 = n.GetS2(quanta.electrons, truthPos[0], truthPos[1], truthPos[2], smearPos[0], smearPos[1],
                                    smearPos[2], driftTime, vD, 0, field, S2CalculationMode::Full, 0, wf_time, wf_amp, g2_params);
    NESTresult full = n.FullCalculation(type_num, keV, rho, field, 131.293, 54., default_NRYieldsParam,
                                        default_NRERWidthsParam, default_ERYieldsParam, false);
    printf("%.6f\t%.6f\t%d\t%d\t%.6f\t%.6f\t%d\t%zu\n", yields.PhotonYield, yields.ElectronYield, quanta.photons,
           quanta.electrons, scint[2], scint2[4], full.quanta.photons + full.quanta.electrons, full.photon_times.size());
    // liquid argon, as examples/LArNEST/LArNESTMeanYieldsBenchmarks.cpp:43-60 calls it
    LArNEST larnest(detector);
    LArNESTResult lar = larnest.FullCalculation(LArInteraction::NR, 10., 0., 500., 1.393, false);
    LArYieldResult lary = larnest.GetYields(LArInteraction::ER, 10., 0., 500., 1.393);
    printf("%.6f\t%.6f\t%.6f\t%.6f\n", lar.yields.Nph, lar.yields.Ne, lar.fluctuations.NphFluctuation, lary.Nph);
    if (quanta.photons + quanta.electrons != quanta.ions + quanta.excitons) status = 2;
    if (full.photon_times.size() != size_t(full.quanta.photons)) status = 2;
  } catch (const exception& e) {
    cerr << "stage_calls: " << e.what() << endl;
    status = 3;
  }
  delete detector;
  return status;
}
