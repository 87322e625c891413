#ifdef FIX_CLASS
// clang-format off
FixStyle(ucgld/langevin, Fix_UCGLD_Langevin);
// clang-format on
#else
#ifndef LMP_FIX_LANGEVIN_UCGLD_H
#define LMP_FIX_LANGEVIN_UCGLD_H

// GPU-backed drop-in for Fix_UCGLD_Langevin (UCG/fix_ucgld_langevin.h:16):
//   fix ID group ucgld/langevin Tstart Tstop period seed
// Thermostat of the lambda degree of freedom; exports "t_target" to the pair styles and to fix ucgstate.

#include "fix.h"
#include "ucg_device.h"

namespace LAMMPS_NS {


class Fix_UCGLD_Langevin : public Fix, public UCGDeckPart {
 protected:
  UCGDevice *dev;
  struct Ramp {            // linear temperature ramp over the run
    double start, stop, period;
    double now, sqrt_now;  // target at the current step
  } temp;
  double *drag, *kick;     // per-type  -m/period/ftm2v  and  sqrt(m) sqrt(24 kB/(period dt mvv2e))/ftm2v
  double *ratio;           // per-type scale factors (fix_modify-compatible, all 1)
  int rng_seed, bias_temp, nlevels_respa;
  char *temp_compute_id;
  class Compute *temperature;
  double lambda_temperature;
  void compute_target();

 public:
  Fix_UCGLD_Langevin(class LAMMPS *, int, char **);
  ~Fix_UCGLD_Langevin() override;
  void init() override;
  void setup(int) override;
  int setmask() override;
  void reset_dt() override;
  void end_of_step() override;
  void post_force(int) override;
  void reset_target(double) override;
  double memory_usage() override;
  double compute_scalar() override;
  int modify_param(int, char **) override;
  void *extract(const char *, int &) override;
  void post_force_respa(int, int, int) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  bool ucg_tracked_ok() const override { return true; }
};

}  // namespace LAMMPS_NS
#endif
#endif
