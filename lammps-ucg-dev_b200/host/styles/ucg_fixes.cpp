// ucg_fixes.cpp — LAMMPS-facing halves of the UCG fixes and of atom_style ucg.  Grammar,
// masks and error texts follow the reference (file:line cited per method); the per-site
// arithmetic runs in csrc/fixes.cu through the C-ABI.
#include "atom_vec_ucg.h"
#include "fix_nve_ucgld.h"
#include "fix_nve_ucgld_wall_hard.h"
#include "fix_ucgld_langevin.h"
#include "fix_ucgstate.h"

#include "atom.h"
#include "comm.h"
#include "compute.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "memory.h"
#include "modify.h"
#include "output.h"
#include "respa.h"
#include "ucg_device.h"
#include "update.h"

#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;
using namespace FixConst;

static const unsigned DYN_IN = UCGB200_F_X | UCGB200_F_V | UCGB200_F_F | UCGB200_F_UCGL | UCGB200_F_UCGVL |
                               UCGB200_F_UCGFORCE | UCGB200_F_UCGSTATE;

// ---------------------------------------------------------------- fix nve/ucgld
// UCG/fix_nve_ucgld.cpp:10-181
FixNVE_UCGLD::FixNVE_UCGLD(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), dt_pos(0), dt_half(0), respa_steps(nullptr), hard_wall(0) {
  if (utils::strmatch(style, "^nve/ucgld$") && narg > 3)
    error->all(FLERR, 3, "Unsupported additional arguments for fix {}", style);
  dynamic_group_allow = 1;
  time_integrate = 1;
  dev = UCGDevice::get(lmp);
}
int FixNVE_UCGLD::setmask() {
  return INITIAL_INTEGRATE | FINAL_INTEGRATE | INITIAL_INTEGRATE_RESPA | FINAL_INTEGRATE_RESPA | PRE_EXCHANGE | END_OF_STEP | POST_RUN;
}
// [stock] Verlet::setup: force evaluation, every Fix::setup (fix ucgstate and fix ucgld/langevin act there), then the
// step-0 thermo line and dumps read the host arrays: set-up stays eager.  From the first initial_integrate to post_run
// the arrays may stay on the device if the deck allows it
void FixNVE_UCGLD::setup(int) { dev->tracked = false; dev->tracking_pending = true; }
// rebuild steps: Domain::pbc, Comm::exchange / borders and Atom::sort are about to rewrite and reorder the host arrays
void FixNVE_UCGLD::pre_exchange() {
  if (!dev->tracked) return;
  dev->flush(lmp);
  dev->host_changed();
}
// output steps: thermo computes and dumps read the host arrays right after end_of_step
void FixNVE_UCGLD::end_of_step() {
  if (dev->tracked && output && output->next == update->ntimestep) dev->flush(lmp);
}
void FixNVE_UCGLD::post_run() { dev->tracking_end(lmp); dev->tracking_pending = false; }

bool UCGDevice::tracking_allowed(LAMMPS *lmp) {
  if (getenv("UCGB200_OFFLOAD_TRACKED") && atoi(getenv("UCGB200_OFFLOAD_TRACKED")) == 0) return false;
  if (lmp->comm->nprocs > 1) return false;
  if (!utils::strmatch(lmp->update->integrate_style, "^verlet")) return false;
  Force *f = lmp->force;
  auto *pp = f->pair ? dynamic_cast<UCGDeckPart *>(f->pair) : nullptr;
  if (!pp || !pp->ucg_tracked_ok()) return false;
  if (f->bond || f->angle || f->dihedral || f->improper || f->kspace) return false;
  Modify *m = lmp->modify;
  for (int i = 0; i < m->nfix; i++) {
    Fix *fx = m->fix[i];
    if (!m->fmask[i]) continue;                                     // does nothing inside the step
    auto *part = dynamic_cast<UCGDeckPart *>(fx);
    if (!part || !part->ucg_tracked_ok()) return false;              // somebody else's fix may read anything
  }
  return true;
}
void FixNVE_UCGLD::init() {
  dt_pos = update->dt;
  dt_half = 0.5 * update->dt * force->ftm2v;
  if (utils::strmatch(update->integrate_style, "^respa")) respa_steps = (dynamic_cast<Respa *>(update->integrate))->step;
  if (atom->rmass) error->all(FLERR, "fix {}: per-atom masses (rmass) are not supported by ucg-b200", style);
}
void FixNVE_UCGLD::initial_integrate(int) {
  if (dev->tracking_pending) { dev->tracking_pending = false; dev->tracking_begin(lmp); }
  dev->upload(lmp, DYN_IN);
  dev->check(lmp, ucgb200_fix_nve_initial(dev->ctx, dt_pos, dt_half, groupbit, hard_wall), "fix_nve_initial");
  dev->download(lmp, UCGB200_F_X | UCGB200_F_V | UCGB200_F_UCGL | UCGB200_F_UCGVL | (hard_wall ? UCGB200_F_UCGSTATE : 0u));
}
void FixNVE_UCGLD::final_integrate() {
  dev->upload(lmp, UCGB200_F_V | UCGB200_F_F | UCGB200_F_UCGVL | UCGB200_F_UCGFORCE | UCGB200_F_UCGL);
  dev->check(lmp, ucgb200_fix_nve_final(dev->ctx, dt_half, groupbit, hard_wall), "fix_nve_final");
  dev->download(lmp, UCGB200_F_V | UCGB200_F_UCGVL | (hard_wall ? UCGB200_F_UCGL : 0u));
}
void FixNVE_UCGLD::initial_integrate_respa(int vflag, int ilevel, int) {
  dt_pos = respa_steps[ilevel];
  dt_half = 0.5 * respa_steps[ilevel] * force->ftm2v;
  if (ilevel == 0) initial_integrate(vflag); else final_integrate();
}
void FixNVE_UCGLD::final_integrate_respa(int ilevel, int) {
  dt_half = 0.5 * respa_steps[ilevel] * force->ftm2v;
  final_integrate();
}
bool FixNVE_UCGLD::ucg_deck(ucgb200_deck &deck) {
  deck.nve = 1;
  deck.nve_groupbit = groupbit;
  return true;
}
void FixNVE_UCGLD::reset_dt() {
  dt_pos = update->dt;
  dt_half = 0.5 * update->dt * force->ftm2v;
}

// ------------------------------------------------------ fix nve/ucgld/hard_wall/hard
// UCG/fix_nve_ucgld_wall_hard.cpp:12-52, 234-257
FixNVE_UCGLD_Wall_Hard::FixNVE_UCGLD_Wall_Hard(LAMMPS *lmp, int narg, char **arg)
    : FixNVE_UCGLD(lmp, narg, arg), bias_on(0), bias_height(0.1) {
  if (narg > 5) error->all(FLERR, 3, "Unsupported additional arguments for fix {}", style);
  hard_wall = 1;
  int iarg = 3;
  while (iarg < narg) {
    if (utils::strmatch(arg[iarg], "bias_potential")) {
      bias_on = 1;
      iarg++;
      if (iarg < narg) bias_height = utils::numeric(FLERR, arg[iarg], false, lmp);
      iarg++;
    } else
      error->all(FLERR, "Unknown argument for fix {}", style);
  }
}
int FixNVE_UCGLD_Wall_Hard::setmask() {
  int mask = FixNVE_UCGLD::setmask();
  if (bias_on) mask |= POST_FORCE;
  return mask;
}
bool FixNVE_UCGLD_Wall_Hard::ucg_deck(ucgb200_deck &deck) {
  deck.nve = 2;
  deck.nve_groupbit = groupbit;
  deck.wall_bias = bias_on;
  deck.wall_barrier = bias_height;
  return true;
}
void FixNVE_UCGLD_Wall_Hard::post_force(int) {
  dev->upload(lmp, UCGB200_F_UCGL | UCGB200_F_UCGFORCE);
  dev->check(lmp, ucgb200_fix_wall_bias(dev->ctx, bias_height, groupbit), "fix_wall_bias");
  dev->download(lmp, UCGB200_F_UCGFORCE);
}

// ----------------------------------------------------------------- fix ucgstate
// UCG/fix_ucgstate.cpp:34-171
FixUCGState::FixUCGState(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), kT_now(0), t_bath(0), lambda_only(0), monte_carlo(0), rng_seed(1), switch_rate(0.01) {
  if (narg > 6) error->all(FLERR, 3, "Too many arguments for fix {}", style);
  if (!atom->ucg_flag) error->all(FLERR, 1, "fix ucgstate requires ucg atom style");
  if (narg > 3) {
    if (utils::strmatch(arg[3], "ld")) lambda_only = 1;
    else if (utils::strmatch(arg[3], "mc")) {
      monte_carlo = 1;
      if (narg == 4) error->all(FLERR, 1, "fix ucgstate mc requires seed and rate information");
      if (narg == 5) error->all(FLERR, 1, "fix ucgstate mc requires rate information");
      rng_seed = utils::inumeric(FLERR, arg[4], false, lmp);
      switch_rate = utils::numeric(FLERR, arg[5], false, lmp);
    } else
      error->all(FLERR, 1, "Unknown argument for fix {}: {}", style, arg[3]);
  }
  dynamic_group_allow = 1;
  time_integrate = 0;
  dev = UCGDevice::get(lmp);
}
int FixUCGState::setmask() { return POST_FORCE | POST_FORCE_RESPA | MIN_POST_FORCE; }
void FixUCGState::post_force(int) {
  dev->upload(lmp, UCGB200_F_SCORES | UCGB200_F_UCGSTATE | UCGB200_F_UCGL);
  const int mode = lambda_only ? 1 : (monte_carlo ? 2 : 0);
  dev->check(lmp, ucgb200_fix_ucgstate(dev->ctx, mode, rng_seed + comm->me, switch_rate, update->ntimestep), "fix_ucgstate");
  dev->download(lmp, UCGB200_F_UCGP | (lambda_only ? 0u : (UCGB200_F_UCGSTATE | UCGB200_F_UCGL)));
}
bool FixUCGState::ucg_deck(ucgb200_deck &deck) {
  deck.ucgstate = lambda_only ? 2 : (monte_carlo ? 3 : 1);
  deck.ucgstate_seed = rng_seed + comm->me;
  deck.ucgstate_rate = switch_rate;
  return true;
}
void FixUCGState::post_force_respa(int vflag, int, int) { post_force(vflag); }
void FixUCGState::min_post_force(int vflag) { post_force(vflag); }
void FixUCGState::setup(int vflag) {
  double *pT = nullptr;
  int pdim;
  for (int ifix = 0; ifix < modify->nfix; ifix++) {
    pT = (double *) modify->fix[ifix]->extract("t_target", pdim);
    if (pT) { t_bath = *pT; break; }
  }
  if (pT == nullptr) error->all(FLERR, "FixUCGState requires a thermostat fix BEFORE ITSELF to set the target temperature t_bath.");
  kT_now = force->boltz * t_bath;
  post_force(vflag);
}

// ------------------------------------------------------------ fix ucgld/langevin
// UCG/fix_ucgld_langevin.cpp:55-417
Fix_UCGLD_Langevin::Fix_UCGLD_Langevin(LAMMPS *lmp, int narg, char **arg)
    : Fix(lmp, narg, arg), drag(nullptr), kick(nullptr), ratio(nullptr), bias_temp(0), nlevels_respa(1),
      temp_compute_id(nullptr), temperature(nullptr), lambda_temperature(0.0) {
  if (narg < 7) error->all(FLERR, "Illegal fix langevin command");
  dynamic_group_allow = 1;
  scalar_flag = 1;
  global_freq = 1;
  extscalar = 1;
  ecouple_flag = 1;
  nevery = 1;
  if (utils::strmatch(arg[3], "^v_")) error->all(FLERR, "lambda dynamic variable is not supported with variable temperature");
  temp.start = utils::numeric(FLERR, arg[3], false, lmp);
  temp.now = temp.start;
  temp.stop = utils::numeric(FLERR, arg[4], false, lmp);
  temp.period = utils::numeric(FLERR, arg[5], false, lmp);
  rng_seed = utils::inumeric(FLERR, arg[6], false, lmp);
  if (temp.period <= 0.0) error->all(FLERR, "Fix langevin period must be > 0.0");
  if (rng_seed <= 0) error->all(FLERR, "Illegal fix langevin command");
  temp.sqrt_now = sqrt(temp.now);
  drag = new double[atom->ntypes + 1];
  kick = new double[atom->ntypes + 1];
  ratio = new double[atom->ntypes + 1];
  for (int i = 0; i <= atom->ntypes; i++) { ratio[i] = 1.0; drag[i] = kick[i] = 0.0; }
  dev = UCGDevice::get(lmp);
}
Fix_UCGLD_Langevin::~Fix_UCGLD_Langevin() {
  if (copymode) return;
  delete[] drag;
  delete[] kick;
  delete[] ratio;
  delete[] temp_compute_id;
}
int Fix_UCGLD_Langevin::setmask() { return POST_FORCE | POST_FORCE_RESPA | END_OF_STEP; }
void Fix_UCGLD_Langevin::init() {
  if (temp_compute_id) {
    temperature = modify->get_compute_by_id(temp_compute_id);
    if (!temperature) error->all(FLERR, "Temperature compute ID {} for fix {} does not exist", temp_compute_id, style);
    if (temperature->tempflag == 0) error->all(FLERR, "Compute ID {} for fix {} does not compute temperature", temp_compute_id, style);
  }
  // (sic) the reference reads atom->ucgml at the TYPE index (:164-171, quirk Q20)
  for (int i = 1; i <= atom->ntypes; i++) {
    drag[i] = -atom->ucgml[i] / temp.period / force->ftm2v;
    kick[i] = sqrt(atom->ucgml[i]) / force->ftm2v;
    kick[i] *= sqrt(24.0 * force->boltz / temp.period / update->dt / force->mvv2e);
    drag[i] *= 1.0 / ratio[i];
    kick[i] *= 1.0 / sqrt(ratio[i]);
  }
  bias_temp = (temperature && temperature->tempbias) ? 1 : 0;
  if (utils::strmatch(update->integrate_style, "^respa")) nlevels_respa = (static_cast<Respa *>(update->integrate))->nlevels;
}
void Fix_UCGLD_Langevin::setup(int vflag) {
  if (utils::strmatch(update->integrate_style, "^verlet")) post_force(vflag);
  else post_force_respa(vflag, nlevels_respa - 1, 0);
}
void Fix_UCGLD_Langevin::compute_target() {
  double delta = update->ntimestep - update->beginstep;
  if (delta != 0.0) delta /= update->endstep - update->beginstep;
  temp.now = temp.start + delta * (temp.stop - temp.start);
  temp.sqrt_now = sqrt(temp.now);
}
void Fix_UCGLD_Langevin::post_force(int) {
  compute_target();
  if (bias_temp) temperature->compute_scalar();
  dev->upload(lmp, UCGB200_F_UCGVL | UCGB200_F_UCGFORCE);
  dev->check(lmp, ucgb200_fix_langevin(dev->ctx, drag, kick, atom->ntypes, temp.sqrt_now, rng_seed + comm->me,
                                       update->ntimestep, groupbit, bias_temp), "fix_langevin");
  dev->download(lmp, UCGB200_F_UCGFORCE);
}
bool Fix_UCGLD_Langevin::ucg_deck(ucgb200_deck &deck) {
  deck.langevin = 1;
  deck.t_start = temp.start;
  deck.t_stop = temp.stop;
  deck.t_period = temp.period;
  deck.langevin_seed = rng_seed + comm->me;
  deck.langevin_groupbit = groupbit;
  // a bias temperature compute changes one rule of the fix (no random force at zero lambda velocity, :285); its
  // compute_scalar() has no effect on the forces (remove_bias / restore_bias are commented out in the reference)
  deck.langevin_bias = bias_temp;
  return true;
}
void Fix_UCGLD_Langevin::post_force_respa(int vflag, int ilevel, int) {
  if (ilevel == nlevels_respa - 1) post_force(vflag);
}
void Fix_UCGLD_Langevin::end_of_step() {
  // (sic) rank-local, divided by nlocal (:303-312)
  dev->upload(lmp, UCGB200_F_UCGVL);
  double ke = 0.0;
  long long cnt = 0;
  dev->check(lmp, ucgb200_lambda_ke(dev->ctx, groupbit, &ke, &cnt), "lambda_ke");
  lambda_temperature = atom->nlocal ? ke / (0.5 * force->boltz * atom->nlocal) : 0.0;
}
void Fix_UCGLD_Langevin::reset_target(double t_new) { temp.now = temp.start = temp.stop = t_new; }
void Fix_UCGLD_Langevin::reset_dt() {
  // (sic) uses atom->mass, not ucgml (:366-376)
  if (atom->mass)
    for (int i = 1; i <= atom->ntypes; i++) {
      kick[i] = sqrt(atom->mass[i]) / force->ftm2v;
      kick[i] *= sqrt(24.0 * force->boltz / temp.period / update->dt / force->mvv2e);
      kick[i] *= 1.0 / sqrt(ratio[i]);
    }
}
int Fix_UCGLD_Langevin::modify_param(int narg, char **arg) {
  if (strcmp(arg[0], "temp") == 0) {
    if (narg < 2) utils::missing_cmd_args(FLERR, "fix_modify", error);
    delete[] temp_compute_id;
    temp_compute_id = utils::strdup(arg[1]);
    temperature = modify->get_compute_by_id(temp_compute_id);
    if (!temperature) error->all(FLERR, "Could not find fix_modify temperature compute ID: {}", temp_compute_id);
    if (temperature->tempflag == 0) error->all(FLERR, "Fix_modify temperature compute {} does not compute temperature", temp_compute_id);
    if (temperature->igroup != igroup && comm->me == 0)
      error->warning(FLERR, "Group for fix_modify temp != fix group: {} vs {}", group->names[igroup], group->names[temperature->igroup]);
    return 2;
  }
  return 0;
}
double Fix_UCGLD_Langevin::compute_scalar() { return lambda_temperature; }
double Fix_UCGLD_Langevin::memory_usage() { return 0.0; }
void *Fix_UCGLD_Langevin::extract(const char *str, int &dim) {
  dim = 0;
  if (strcmp(str, "t_target") == 0) return &temp.now;
  return nullptr;
}

// --------------------------------------------------------------- atom_style ucg
// UCG/atom_vec_ucg.cpp:29-234 (the field lists define every message the halo kernels carry)
AtomVecUCG::AtomVecUCG(LAMMPS *lmp) : AtomVec(lmp) {
  molecular = Atom::MOLECULAR;
  bonds_allow = angles_allow = dihedrals_allow = impropers_allow = 1;
  mass_type = PER_TYPE;
  forceclearflag = 1;
  atom->molecule_flag = atom->q_flag = 1;
  atom->ucg_flag = 1;
  atom->max_ucgstates = 2;
  const std::vector<std::string> topo = {"q", "molecule", "num_bond", "bond_type", "bond_atom", "num_angle", "angle_type",
      "angle_atom1", "angle_atom2", "angle_atom3", "num_dihedral", "dihedral_type", "dihedral_atom1", "dihedral_atom2",
      "dihedral_atom3", "dihedral_atom4", "num_improper", "improper_type", "improper_atom1", "improper_atom2",
      "improper_atom3", "improper_atom4", "nspecial", "special"};
  const std::vector<std::string> ucg = {"ucgstate", "ucgl", "ucgvl", "ucgml", "ucgp", "ucgforce", "ucgsoftmaxscores", "num_ucgstates"};
  fields_grow = topo; fields_grow.insert(fields_grow.end(), ucg.begin(), ucg.end());
  fields_copy = fields_grow;
  fields_exchange = fields_grow;
  fields_border = {"q", "molecule", "ucgstate", "num_ucgstates", "ucgl", "ucgp"};
  fields_border_vel = {"q", "molecule", "ucgstate", "num_ucgstates", "ucgl", "ucgp", "ucgvl"};
  fields_comm = {"ucgstate", "ucgl", "ucgp"};
  fields_comm_vel = {"ucgstate", "ucgl", "ucgvl", "ucgp"};
  fields_reverse = {"ucgforce", "ucgsoftmaxscores"};
  fields_restart = {"ucgstate", "ucgl", "ucgml", "ucgvl", "ucgp"};
  fields_data_atom = {"id", "molecule", "type", "q", "x", "ucgstate", "ucgl", "ucgml"};
  fields_data_vel = {"id", "v", "ucgvl"};
  setup_fields();
}
void AtomVecUCG::grow_pointers() {
  topo.bonds = atom->num_bond; topo.angles = atom->num_angle; topo.dihedrals = atom->num_dihedral; topo.impropers = atom->num_improper;
  topo.special_counts = atom->nspecial;
  site.state = atom->ucgstate; site.lambda = atom->ucgl; site.flambda = atom->ucgforce; site.scores = atom->ucgsoftmaxscores;
  site.vlambda = atom->ucgvl; site.prob = atom->ucgp; site.mlambda = atom->ucgml; site.nstates = atom->num_ucgstates;
}
void AtomVecUCG::force_clear(int n, size_t nbytes) {
  memset(&site.flambda[n], 0, nbytes);
  memset(&site.scores[n][0], 0, atom->max_ucgstates * nbytes);
}
void AtomVecUCG::data_atom_post(int ilocal) {
  topo.bonds[ilocal] = topo.angles[ilocal] = topo.dihedrals[ilocal] = topo.impropers[ilocal] = 0;
  topo.special_counts[ilocal][0] = topo.special_counts[ilocal][1] = topo.special_counts[ilocal][2] = 0;
  if (site.lambda[ilocal] < 0) site.lambda[ilocal] = 0.;
  else if (site.lambda[ilocal] > 1) site.lambda[ilocal] = 1.;
  if (site.state[ilocal] < 0) site.state[ilocal] = 0;
  else if (site.state[ilocal] > 1) site.state[ilocal] = 1;
  site.prob[ilocal] = -1.0;   // "unassigned" until the first force evaluation
}
int AtomVecUCG::property_atom(const std::string &name) {
  static const char *names[] = {"ucgstate", "ucgl", "ucgforce", "ucgvl", "ucgp", "ucgml"};
  for (int k = 0; k < 6; k++) if (name == names[k]) return k;
  return -1;
}
void AtomVecUCG::pack_property_atom(int index, double *buf, int nvalues, int groupbit) {
  if (index < 0 || index > 5) error->all(FLERR, "Unknown property_atom index in AtomVecUCG::pack_property_atom");
  const int *mask = atom->mask;
  const int nlocal = atom->nlocal;
  const double *src[6] = {nullptr, site.lambda, site.flambda, site.vlambda, site.prob, site.mlambda};
  int n = 0;
  for (int j = 0; j < nlocal; j++, n += nvalues) {
    if (!(mask[j] & groupbit)) buf[n] = 0.0;
    else buf[n] = index == 0 ? (double) site.state[j] : src[index][j];
  }
}
