#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_rleucg_interface, PairTable_RLEUCG_INTERFACE)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_RLEUCG_INTERFACE_H
#define LMP_PAIR_TABLE_RLEUCG_INTERFACE_H

// Registered name and method set of the reference's PairTable_RLEUCG_INTERFACE; the device kernels for this
// style are not built yet: every entry point fails loudly (there is no CPU fallback).

#include "pair.h"

namespace LAMMPS_NS {

class PairTable_RLEUCG_INTERFACE : public Pair {
 public:
  PairTable_RLEUCG_INTERFACE(class LAMMPS *lmp) : Pair(lmp) {}
  void compute(int, int) override { error->all(FLERR, "pair_style table_rleucg_interface: sm_100a kernels not built in this release of ucg-b200"); }
  void settings(int, char **) override { error->all(FLERR, "pair_style table_rleucg_interface: sm_100a kernels not built in this release of ucg-b200"); }
  void coeff(int, char **) override { error->all(FLERR, "pair_style table_rleucg_interface: sm_100a kernels not built in this release of ucg-b200"); }
};

}  // namespace LAMMPS_NS
#endif
#endif
