/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the reference's hot path,
 * written from the reference's algorithm (file:line cited at each function in
 * nest_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (nest_b200/) never includes, links or calls it.
 *
 * Parity pin: with the same seed this library reproduces the UNMODIFIED reference
 * (oracle/_ref, built from /root/reference by oracle/Makefile) bit for bit --
 * same xoroshiro128+ stream, same draw order, same libm calls, and libstdc++'s own
 * std::binomial_distribution / std::poisson_distribution -- checked by
 * tests/test_oracle_pin.py against the reference binary's stdout and against
 * fixtures in tests/golden/ generated from it (tests/golden/make_golden.py).
 * The reference itself ships no golden vectors for this path.
 */
#ifndef NEST_ORACLE_H
#define NEST_ORACLE_H 1
#include <stdint.h>

#include "../include/nest_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* columns of orc_chain output, AoS [n][ORC_NCOL] (same layout as ref_probe.cpp) */
enum {
  ORC_KEV_TRUE = 0, ORC_KEV_OUT, ORC_FIELD, ORC_DT, ORC_X, ORC_Y, ORC_Z, ORC_SX, ORC_SY,
  ORC_NPH, ORC_NE, ORC_NI, ORC_NEX, ORC_YPH, ORC_YNE, ORC_S1_0,
  ORC_S2_0 = ORC_S1_0 + 9, ORC_SIG1 = ORC_S2_0 + 9, ORC_SIG2, ORC_SIGE, ORC_NCOL
};

int orc_chain_ncol(void);
double orc_density(double T, double P, int* inGas);
void orc_work_function(double rho, double molarMass, int OldW13eV, double* Wq_eV, double* alpha);
double orc_drift_velocity_liquid(double T, double F);
double orc_find_new_mean(double sigma);
int orc_prepare(const nb200_detector* det, double inField, nb200_run_consts* out);
void orc_maps(const nb200_detector* det, int n, const double* x, const double* y, const double* z,
              double* s1, double* s2, double* ef, double* tba);
int orc_yields(const nb200_detector* det, int which, int species, int n, const double* E,
               const double* field, double rho, double massNum, double atomNum, const double* nrpar,
               const double* erpar, double multFact, double* out);
void orc_widths(const nb200_detector* det, int n, const double* efield, const double* elecFrac,
                const double* nq, double rho, const double* widths, int nw, double* out);
int orc_quanta(const nb200_detector* det, uint64_t seed, int n, const double* y6, double rho,
               const double* widths, int nw, double skewER, int* iout, double* dout);
void orc_sampler(int kind, uint64_t seed, int n, double p0, double p1, double p2, double* out);
void orc_spectrum(int kind, uint64_t seed, int n, double eMin, double eMax, double* out);
void orc_s1(const nb200_detector* det, uint64_t seed, int n, int Nph, const double* pos, const double* spos,
            double vD, double vDmid, int species, double field, double energy, int mode, double* out);
void orc_s2(const nb200_detector* det, uint64_t seed, int n, int Ne, const double* pos, const double* spos,
            double dt, double vD, double field, double* out);
void orc_xyres(const nb200_detector* det, uint64_t seed, int n, double x, double y, double A_top, double* out);
int orc_chain(const nb200_detector* det, const nb200_analysis* ana, uint64_t seed, int n, int species,
              int src_kind, double eMin, double eMax, double inField, int fixed_pos, const double* xyz,
              int er_model, double multFact, const double* nrpar, const double* erpar,
              const double* widths, int nw, double skewER, int g2_verbosity, double* out);
int orc_getband(const nb200_analysis* ana, int coinLevel, int n, const double* s1, const double* s2,
                double* out);
int orc_runvec(const nb200_detector* det, int species, int n, const double* E, const double* xyz,
               double inField, int seed, const double* erpar, const double* nrpar, const double* widths,
               int nw, int s1mode, double* out);
void orc_lar(uint64_t seed, int species, int n, const double* E, const double* field, double density,
             double* yields, double* fluct);
#ifdef __cplusplus
}
#endif
#endif
