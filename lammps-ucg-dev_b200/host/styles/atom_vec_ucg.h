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
 protected:
  struct Topology {        // per-atom topology counters reset in data_atom_post
    int *bonds, *angles, *dihedrals, *impropers;
    int **special_counts;
  } topo;
  struct Site {            // cached atom->ucg* pointers (re-taken in grow_pointers)
    int *state, *nstates;
    double *lambda, *vlambda, *mlambda, *prob, *flambda;
    double **scores;
  } site;

 public:
  AtomVecUCG(class LAMMPS *);
  void grow_pointers() override;
  void data_atom_post(int) override;
  void force_clear(int, size_t) override;
  int property_atom(const std::string &name) override;
  void pack_property_atom(int, double *, int, int) override;
};

}  // namespace LAMMPS_NS
#endif
#endif
