// PairTable_UCG_Bethe on the GPU: keyword grammar of UCG/pair_table_ucg_bethe.cpp:746-868.
#include "pair_table_ucg_bethe.h"

#include "error.h"
#include "ucg_device.h"

#include <cstring>
#include <vector>

using namespace LAMMPS_NS;

PairTable_UCG_Bethe::PairTable_UCG_Bethe(LAMMPS *lmp)
    : PairTable_UCGLD(lmp), method_flag(BETHE), pseudo_flag(0), prior_flag(CHEMICAL_POTENTIAL), seed(1), noise_level(0.0) {}

void PairTable_UCG_Bethe::settings(int narg, char **arg) {
  std::vector<char *> rest;
  for (int iarg = 0; iarg < narg; iarg++) {
    if (iarg >= 3 && strcmp(arg[iarg], "method") == 0) {
      if (++iarg >= narg) error->all(FLERR, "Missing argument for pair_style table_ucg_bethe method");
      if (!strcmp(arg[iarg], "mf") || !strcmp(arg[iarg], "meanfield")) method_flag = MF;
      else if (!strcmp(arg[iarg], "bethe") || !strcmp(arg[iarg], "Bethe")) method_flag = BETHE;
      else error->all(FLERR, "Unknown argument for pair_style table_ucg_bethe method: {}, please write mf or bethe", arg[iarg]);
    } else if (iarg >= 3 && strcmp(arg[iarg], "pseudo") == 0) {
      if (++iarg >= narg) error->all(FLERR, "Missing argument for pair_style table_ucg_bethe pseudo");
      // (sic) "yes" -> flag 0, "no" -> flag 1, as in the reference (:826-837)
      if (!strcmp(arg[iarg], "yes")) pseudo_flag = 0;
      else if (!strcmp(arg[iarg], "no")) pseudo_flag = 1;
      else error->all(FLERR, "Unknown argument for pair_style table_ucg_bethe pseudo: {}, please write yes or no", arg[iarg]);
    } else if (iarg >= 3 && strcmp(arg[iarg], "prior") == 0) {
      if (++iarg >= narg) error->all(FLERR, "Missing argument for pair_style table_ucg_bethe");
      if (!strcmp(arg[iarg], "chemical_potential")) {
        prior_flag = CHEMICAL_POTENTIAL;
        if (iarg + 1 < narg && !strcmp(arg[iarg + 1], "noise")) {
          prior_flag = CHEMICAL_POTENTIAL_NOISE;
          if (iarg + 3 >= narg) error->all(FLERR, "Missing argument for pair_style table_ucg_bethe prior chemical_potential noise");
          noise_level = utils::numeric(FLERR, arg[iarg + 2], false, lmp);
          if (noise_level <= 0.0) noise_level = 0.0;
          seed = utils::inumeric(FLERR, arg[iarg + 3], false, lmp);
          if (seed <= 0) seed = -seed + 1;
          iarg += 3;
        }
      } else if (!strcmp(arg[iarg], "ucgl")) prior_flag = UCGL;
      else error->all(FLERR, "Unknown argument for pair_style table_ucg_bethe prior: {}, please write chemical_potential or ucgl", arg[iarg]);
    } else
      rest.push_back(arg[iarg]);
  }
  PairTable_UCGLD::settings((int) rest.size(), rest.data());
}

bool PairTable_UCG_Bethe::ucg_deck(ucgb200_deck &deck) {
  PairTable_UCGLD::ucg_deck(deck);
  deck.pair_style = 1;
  deck.bethe_method = method_flag;
  deck.bethe_pseudo = pseudo_flag;
  deck.bethe_prior = prior_flag;
  deck.bethe_noise_level = noise_level;   // acts on the first evaluation only (ucgp still -1): the same draws as in offload mode
  deck.bethe_seed = seed;
  return true;
}

void PairTable_UCG_Bethe::device_compute(int eflag, int vflag) {
  dev->check(lmp, ucgb200_pair_bethe(dev->ctx, eflag, vflag, method_flag, pseudo_flag, prior_flag, noise_level, seed), "pair_bethe");
}
