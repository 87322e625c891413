#ifdef FIX_CLASS
// clang-format off
FixStyle(nve/ucgld/wall/hard, FixNVE_UCGLD_Wall_Hard);
// clang-format on
#else
#ifndef LMP_FIX_NVE_UCGLD_WALL_HARD_H
#define LMP_FIX_NVE_UCGLD_WALL_HARD_H

// GPU-backed drop-in for FixNVE_UCGLD_Wall_Hard (UCG/fix_nve_ucgld_wall_hard.h:16):
//   fix ID group nve/ucgld/wall/hard [bias_potential H]
// The reflection lives in the integrator kernels of the base class; this class adds the optional
// double-well bias on lambda as a post_force stage.

#include "fix_nve_ucgld.h"

namespace LAMMPS_NS {

class FixNVE_UCGLD_Wall_Hard : public FixNVE_UCGLD {
 protected:
  double bias_height;     // H of  (-7980 x^9 + 2 x) * 10 H,  x = lambda - 1/2
  int bias_on;            // keyword bias_potential given

 public:
  FixNVE_UCGLD_Wall_Hard(class LAMMPS *, int, char **);
  void post_force(int) override;
  int setmask() override;
  bool ucg_deck(ucgb200_deck &deck) override;
};

}  // namespace LAMMPS_NS
#endif
#endif
