#ifdef FIX_CLASS
// clang-format off
FixStyle(ucgstate, FixUCGState);
// clang-format on
#else
#ifndef LMP_FIX_UCGSTATE_H
#define LMP_FIX_UCGSTATE_H

// GPU-backed drop-in for FixUCGState (UCG/fix_ucgstate.h:3):   fix ID group ucgstate [ld | mc seed rate]
// softmax of the per-site scores -> ucgp; without `ld` also the discrete state (rounded, or Monte-Carlo
// with a counter-based RNG keyed by site tag and time step) and ucgl = ucgp.

#include "fix.h"
#include "ucg_device.h"

namespace LAMMPS_NS {


class FixUCGState : public Fix, public UCGDeckPart {
  UCGDevice *dev;
  int lambda_only;        // `ld`: probabilities only, lambda dynamics owns the state
  int monte_carlo;        // `mc seed rate`
  int rng_seed;
  double switch_rate;
  double t_bath, kT_now;  // from the first fix that exports "t_target"

 public:
  FixUCGState(class LAMMPS *, int, char **);
  void setup(int) override;
  int setmask() override;
  void post_force(int) override;
  void min_post_force(int) override;
  void post_force_respa(int, int, int) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  bool ucg_tracked_ok() const override { return true; }
};

}  // namespace LAMMPS_NS
#endif
#endif
