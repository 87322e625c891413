#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_ucg_bethe_density, PairTable_UCG_Bethe_Density)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_UCG_BETHE_DENSITY_H
#define LMP_PAIR_TABLE_UCG_BETHE_DENSITY_H

// Registered name and method set of the reference's PairTable_UCG_Bethe_Density; the device kernels for this
// style are not built yet: every entry point fails loudly (there is no CPU fallback).

#include "pair.h"

namespace LAMMPS_NS {

class PairTable_UCG_Bethe_Density : public Pair {
 public:
  PairTable_UCG_Bethe_Density(class LAMMPS *lmp) : Pair(lmp) {}
  void compute(int, int) override { error->all(FLERR, "pair_style table_ucg_bethe_density: sm_100a kernels not built in this release of ucg-b200"); }
  void settings(int, char **) override { error->all(FLERR, "pair_style table_ucg_bethe_density: sm_100a kernels not built in this release of ucg-b200"); }
  void coeff(int, char **) override { error->all(FLERR, "pair_style table_ucg_bethe_density: sm_100a kernels not built in this release of ucg-b200"); }
};

}  // namespace LAMMPS_NS
#endif
#endif
