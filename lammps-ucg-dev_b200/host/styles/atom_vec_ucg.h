#ifdef ATOM_CLASS
// clang-format off
AtomStyle(ucg, AtomVecUCG);
// clang-format on
#else
#ifndef LMP_ATOM_VEC_UCG_H
#define LMP_ATOM_VEC_UCG_H

// atom_style ucg (UCG/atom_vec_ucg.h:22): `full` + the UCG per-site arrays.  The host-side
// schema (field lists -> every MPI message, data-file columns, clamps) is what the device
// records of libucgb200 mirror; see DESIGN.md §3.

#include "atom_vec.h"

namespace LAMMPS_NS {

class AtomVecUCG : virtual public AtomVec {
 public:
  AtomVecUCG(class LAMMPS *);
  void grow_pointers() override;
  void force_clear(int, size_t) override;
  void data_atom_post(int) override;
  int property_atom(const std::string &name) override;
  void pack_property_atom(int, double *, int, int) override;

 protected:
  int *num_bond, *num_angle, *num_dihedral, *num_improper;
  int **nspecial;
  int *ucgstate, *num_ucgstates;
  double *ucgl, *ucgvl, *ucgml, *ucgp, *ucgforce;
  double **ucgsoftmaxscores;
};

}  // namespace LAMMPS_NS
#endif
#endif
