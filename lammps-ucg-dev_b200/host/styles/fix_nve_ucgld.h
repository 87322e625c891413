#ifdef FIX_CLASS
// clang-format off
FixStyle(nve/ucgld, FixNVE_UCGLD);
// clang-format on
#else
#ifndef LMP_FIX_NVE_UCGLD_H
#define LMP_FIX_NVE_UCGLD_H

// GPU-backed drop-in for FixNVE_UCGLD (UCG/fix_nve_ucgld.h:16): velocity Verlet for the
// particles and for lambda.

#include "fix.h"

namespace LAMMPS_NS {

class FixNVE_UCGLD : public Fix {
 public:
  FixNVE_UCGLD(class LAMMPS *, int, char **);
  int setmask() override;
  void init() override;
  void initial_integrate(int) override;
  void final_integrate() override;
  void initial_integrate_respa(int, int, int) override;
  void final_integrate_respa(int, int) override;
  void reset_dt() override;

 protected:
  double dtv, dtf;
  double *step_respa;
  int wall;               // 1 in the wall/hard subclass
  class UCGDevice *dev;
};

}  // namespace LAMMPS_NS
#endif
#endif
