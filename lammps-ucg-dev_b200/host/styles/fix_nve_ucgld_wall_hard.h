#ifdef FIX_CLASS
// clang-format off
FixStyle(nve/ucgld/wall/hard, FixNVE_UCGLD_Wall_Hard);
// clang-format on
#else
#ifndef LMP_FIX_NVE_UCGLD_WALL_HARD_H
#define LMP_FIX_NVE_UCGLD_WALL_HARD_H

// GPU-backed drop-in for FixNVE_UCGLD_Wall_Hard (UCG/fix_nve_ucgld_wall_hard.h:16):
// fix ID group nve/ucgld/wall/hard [bias_potential H]

#include "fix_nve_ucgld.h"

namespace LAMMPS_NS {

class FixNVE_UCGLD_Wall_Hard : public FixNVE_UCGLD {
 public:
  FixNVE_UCGLD_Wall_Hard(class LAMMPS *, int, char **);
  int setmask() override;
  void post_force(int) override;

 protected:
  int bias_potential_flag;
  double barrier;
};

}  // namespace LAMMPS_NS
#endif
#endif
