#ifdef PAIR_CLASS
// clang-format off
PairStyle(table_rleucg_interface, PairTable_RLEUCG_INTERFACE)
// clang-format on
#else
#ifndef LMP_PAIR_TABLE_RLEUCG_INTERFACE_H
#define LMP_PAIR_TABLE_RLEUCG_INTERFACE_H

// GPU-backed drop-in for PairTable_RLEUCG_INTERFACE (UCG/pair_table_rleucg_interface.h:16-110):
//   pair_style table_rleucg_interface <style> <N> <statefile>
//   pair_coeff i j <file> <keyword> [cut]            (i, j are STATE types, as in pair_style table)
// State file (read_state_settings, .cpp:575-664):
//   n_actual n_total_states
//   per actual type:  <n_states> density use_entropy|no_entropy
//     if > 1 state:   <density threshold> <threshold radius>
//                     <mu_1 ... mu_{n-1}>
// Full neighbor list, newton off (init_style :759-797).

#include "pair.h"
#include "ucgb200_host.h"
#include "ucg_device.h"

#include <vector>

namespace LAMMPS_NS {

class PairTable_RLEUCG_INTERFACE : public Pair, public UCGDeckPart {
 public:
  PairTable_RLEUCG_INTERFACE(class LAMMPS *);
  ~PairTable_RLEUCG_INTERFACE() override;
  void compute(int, int) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;
  void write_restart(FILE *) override;
  void read_restart(FILE *) override;
  void write_restart_settings(FILE *) override;
  void read_restart_settings(FILE *) override;
  double single(int, int, int, int, double, double, double, double &) override;
  void *extract(const char *, int &) override;
  // the ghost exchanges of (p, dp/drho, F_p) and of the CV back-force happen on the device
  // (pair_table_rleucg_interface.cpp:102-160 over MPI); nothing is packed on the host
  int pack_forward_comm(int, int *, double *, int, int *) override { return 0; }
  void unpack_forward_comm(int, int, double *) override {}
  int pack_reverse_comm(int, int, double *) override { return 0; }
  void unpack_reverse_comm(int, int *, double *) override {}
  enum { LOOKUP, LINEAR, SPLINE, BITMAP };

 protected:
  int tabstyle, tablength;
  double T, kT;
  int n_actual_types, n_total_states;
  std::vector<int> n_states_per_type, actual_types_from_state, use_state_entropy;
  std::vector<double> chemical_potentials, cv_thresholds, threshold_radii;
  std::vector<ucgb200_table *> tables;
  std::vector<double> tabcut;
  int **tabindex;
  bool configured;
  class UCGDevice *dev;

  void allocate();
  void read_state_settings(const char *);
  void configure_device();
};

}  // namespace LAMMPS_NS
#endif
#endif
