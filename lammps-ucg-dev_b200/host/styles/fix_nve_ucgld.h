#ifdef FIX_CLASS
// clang-format off
FixStyle(nve/ucgld, FixNVE_UCGLD);
// clang-format on
#else
#ifndef LMP_FIX_NVE_UCGLD_H
#define LMP_FIX_NVE_UCGLD_H

// GPU-backed drop-in for FixNVE_UCGLD (UCG/fix_nve_ucgld.h:16): velocity Verlet for the particles and for
// lambda; both half-steps are kernels of csrc/fixes.cu.  The hard-wall variant derives from this class and
// only flips `hard_wall`.

#include "fix.h"
#include "ucg_device.h"

namespace LAMMPS_NS {


class FixNVE_UCGLD : public Fix, public UCGDeckPart {
 protected:
  UCGDevice *dev;         // the device context shared by every UCG style of this LAMMPS instance
  int hard_wall;          // 1: lambda is reflected at 0 and 1 and the discrete state follows lambda
  double dt_pos;          // update->dt
  double dt_half;         // 0.5 * update->dt * force->ftm2v
  double *respa_steps;    // Respa::step when run_style respa is active

 public:
  FixNVE_UCGLD(class LAMMPS *, int, char **);
  void init() override;
  int setmask() override;
  void reset_dt() override;
  void final_integrate() override;
  void initial_integrate(int) override;
  void final_integrate_respa(int, int) override;
  void initial_integrate_respa(int, int, int) override;
  bool ucg_deck(ucgb200_deck &deck) override;
  bool ucg_tracked_ok() const override { return true; }
  // tracked offload mode (ucg_device.h): the integrator fix opens the window in which per-atom arrays stay on the
  // device, and hands them back where stock LAMMPS reads or reorders them
  void setup(int) override;
  void pre_exchange() override;
  void end_of_step() override;
  void post_run() override;
};

}  // namespace LAMMPS_NS
#endif
#endif
