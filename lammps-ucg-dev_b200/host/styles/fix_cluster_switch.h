#ifdef FIX_CLASS
// clang-format off
FixStyle(cluster_switch, FixClusterSwitch);
// clang-format on
#else
#ifndef LMP_FIX_CLUSTER_SWITCH_H
#define LMP_FIX_CLUSTER_SWITCH_H

// Registered name of the reference's FixClusterSwitch (UCG/fix_cluster_switch.h:3); the device
// kernels for it are not built yet: the constructor fails loudly (no CPU fallback).

#include "fix.h"

namespace LAMMPS_NS {

class FixClusterSwitch : public Fix {
 public:
  FixClusterSwitch(class LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg) {
    error->all(FLERR, "fix cluster_switch: sm_100a kernels not built in this release of ucg-b200");
  }
  int setmask() override { return 0; }
};

}  // namespace LAMMPS_NS
#endif
#endif
