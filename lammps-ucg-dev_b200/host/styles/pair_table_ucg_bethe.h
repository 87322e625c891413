#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_ucg_bethe, PairTable_UCG_Bethe)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_UCG_BETHE_H
#define LMP_PAIR_TABLE_UCG_BETHE_H

// GPU-backed drop-in for PairTable_UCG_Bethe (UCG/pair_table_ucg_bethe.h:33):
// pair_style table_ucg_bethe <style> <N> <statefile> [method mf|bethe] [pseudo yes|no]
//                            [prior chemical_potential [noise lvl seed] | ucgl]

#include "pair_table_ucgld.h"

namespace LAMMPS_NS {

class PairTable_UCG_Bethe : public PairTable_UCGLD {
 public:
  PairTable_UCG_Bethe(class LAMMPS *);
  void settings(int, char **) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  enum { MF, BETHE };
  enum { CHEMICAL_POTENTIAL, CHEMICAL_POTENTIAL_NOISE, UCGL };

 protected:
  int method_flag, pseudo_flag, prior_flag, seed;
  double noise_level;
  void device_compute(int eflag, int vflag) override;
};

}  // namespace LAMMPS_NS
#endif
#endif
