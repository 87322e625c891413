// ucg_host.h — host-side set-up objects shared by the LAMMPS style classes and the
// C-ABI helpers in include/ucgb200_host.h.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ucgb200_host.h"

namespace ucg_host {

// one section of a LAMMPS table file, as read_table() leaves it (pair_table_ucgld.cpp:897-1017)
struct TableInput {
  int ninput = 0, rflag = UCGB200_R_NONE, fpflag = 0;
  double rlo = 0, rhi = 0, fplo = 0, fphi = 0;
  std::vector<double> r, e, f;
  static TableInput from_file(const std::string &file, const std::string &keyword);
  void regenerate_r();
};

// the run-time table compute_table() produces (pair_table_ucgld.cpp:1105-1344)
struct BuiltTable {
  int style = UCGB200_TAB_LINEAR, tablength = 0, match = 0, nmask = 0, nshiftbits = 0;
  double innersq = 0, delta = 0, invdelta = 0, deltasq6 = 0, cut = 0;
  std::vector<double> rsq, drsq, e, de, f, df, e2, f2;
  static BuiltTable build(TableInput in, double cut, int tabstyle, int tablength);
  int single(double rsq, double factor_lj, double &phi, double &fforce) const;
};

// type maps of read_state_settings + the tabindex/setflag/cutsq bookkeeping of coeff/init_one
struct StateMap {
  int n_actual = 0, n_formal = 0;
  std::vector<int> n_states, formal_from, tabindex, setflag;
  std::vector<double> chem_pot, cutsq, tabcut;
  static StateMap from_file(const std::string &file);
  void allocate();
  void coeff(int ilo, int ihi, int jlo, int jhi, int ns_i, int ns_j, const int *tables, const double *cuts);
  void init();
};

}  // namespace ucg_host
