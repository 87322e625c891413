#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_ucg_bethe_density, PairTable_UCG_Bethe_Density)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_UCG_BETHE_DENSITY_H
#define LMP_PAIR_TABLE_UCG_BETHE_DENSITY_H

// GPU-backed drop-in for PairTable_UCG_Bethe_Density (UCG/pair_table_ucg_bethe_density.h:29-110):
//   pair_style table_ucg_bethe_density <style> <N> <statefile>
//   pair_coeff i j Ns_i Ns_j <file keyword cut> x Ns_i*Ns_j          (as table_ucgld)
// State file (read_state_settings, .cpp:778-893):
//   n_actual n_formal max_states
//   per actual type:  <type> <n_states>
//     if 2 states:    <formal0> <formal1> density|<other> entropy|no_entropy
//                     <density threshold> <threshold radius>        (only with "density")
//                     <mu0> <mu1>
// Full neighbor list, newton off (init_style :1103-1153).  Semantics are the repaired ones
// (SURVEY Q9-Q12, Q14); the shipped file cannot get past pair_coeff.

#include "pair_table_ucgld.h"

namespace LAMMPS_NS {

class PairTable_UCG_Bethe_Density : public PairTable_UCGLD {
 public:
  PairTable_UCG_Bethe_Density(class LAMMPS *);
  void compute(int, int) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  void configure_device();
  void settings(int, char **) override;
  void init_style() override;

 protected:
  std::vector<int> use_density, use_state_entropy;
  std::vector<double> cv_thresholds, threshold_radii;
  bool density_applied;
  void read_state_settings(const char *);
};

}  // namespace LAMMPS_NS
#endif
#endif
