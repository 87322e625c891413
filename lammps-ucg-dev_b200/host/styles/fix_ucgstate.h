#ifdef FIX_CLASS
// clang-format off
FixStyle(ucgstate, FixUCGState);
// clang-format on
#else
#ifndef LMP_FIX_UCGSTATE_H
#define LMP_FIX_UCGSTATE_H

// GPU-backed drop-in for FixUCGState (UCG/fix_ucgstate.h:3): fix ID group ucgstate [ld | mc seed rate]

#include "fix.h"

namespace LAMMPS_NS {

class FixUCGState : public Fix {
 public:
  FixUCGState(class LAMMPS *, int, char **);
  int setmask() override;
  void post_force(int) override;
  void post_force_respa(int, int, int) override;
  void min_post_force(int) override;
  void setup(int) override;

 private:
  double kT, T;
  int ld_flag, mc_flag, mc_seed;
  double mc_rate;
  class UCGDevice *dev;
};

}  // namespace LAMMPS_NS
#endif
#endif
