#ifdef FIX_CLASS
// clang-format off
FixStyle(ucgld/langevin, Fix_UCGLD_Langevin);
// clang-format on
#else
#ifndef LMP_FIX_LANGEVIN_UCGLD_H
#define LMP_FIX_LANGEVIN_UCGLD_H

// GPU-backed drop-in for Fix_UCGLD_Langevin (UCG/fix_ucgld_langevin.h:16):
// fix ID group ucgld/langevin Tstart Tstop period seed

#include "fix.h"

namespace LAMMPS_NS {

class Fix_UCGLD_Langevin : public Fix {
 public:
  Fix_UCGLD_Langevin(class LAMMPS *, int, char **);
  ~Fix_UCGLD_Langevin() override;
  int setmask() override;
  void init() override;
  void setup(int) override;
  void post_force(int) override;
  void post_force_respa(int, int, int) override;
  void end_of_step() override;
  void reset_target(double) override;
  void reset_dt() override;
  int modify_param(int, char **) override;
  double compute_scalar() override;
  double memory_usage() override;
  void *extract(const char *, int &) override;

 protected:
  double t_start, t_stop, t_period, t_target, tsqrt;
  double *gfactor1, *gfactor2, *ratio;
  int seed, tbiasflag, nlevels_respa;
  char *id_temp;
  class Compute *temperature;
  double lambda_temp;
  class UCGDevice *dev;
  void compute_target();
};

}  // namespace LAMMPS_NS
#endif
#endif
