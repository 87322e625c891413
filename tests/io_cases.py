"""Shared definitions of the dump / read_dump parity cases: the deterministic per-site state, the dump decks
and the read_dump decks.  Used by tests/golden/make_golden_io.py (which runs them through the reference's own
dump_custom.cpp / read_dump.cpp in oracle/_ref and stores the results under tests/golden/io/), by the oracle
pin tests and by the GPU parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")
GROUP_BIT = 2          # group "half": every site whose id is divisible by 3 carries bit 2 as well
MASS = np.array([0.0, 1.5, 2.5])


def make_state(synth, ncell=3, seed=7, shuffle=True):
    """a small UCG liquid with every dumped quantity non-trivial: forces over 16 decades (all three %g
    notations), two types, two groups, molecule ids, ids shuffled against the storage order"""
    liq = synth.fcc_liquid(ncell, mol_size=4)
    n = liq.n
    rng = np.random.default_rng(seed)
    if shuffle:
        perm = rng.permutation(n)
        for k in ("x", "v", "type", "mask", "tag", "molecule", "ucgstate", "ucgl", "ucgvl", "ucgml"):
            setattr(liq, k, np.ascontiguousarray(getattr(liq, k)[perm]))
    liq.type = (1 + (liq.tag % 5 == 0)).astype(np.int32)
    liq.mask = (1 | np.where(liq.tag % 3 == 0, GROUP_BIT, 0)).astype(np.int32)
    f = rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-8, 9, size=(n, 3))
    f[rng.integers(0, n, 5)] = 0.0
    dyn = dict(f=f, ucgforce=rng.normal(size=n) * 3.0, ucgp=np.clip(rng.uniform(-0.1, 1.1, n), 1e-6, 1 - 1e-6),
               scores=rng.normal(size=(n, 2)))
    liq.ucgml = rng.choice([5.0, 10.0, 12.5], n)
    return liq, dyn


def atoms_dict(liq, dyn):
    return dict(x=liq.x.copy(), v=liq.v.copy(), f=dyn["f"].copy(), type=liq.type.copy(), mask=liq.mask.copy(), tag=liq.tag.copy(),
                molecule=liq.molecule.copy(), ucgstate=liq.ucgstate.copy(), ucgl=liq.ucgl.copy(), ucgvl=liq.ucgvl.copy(),
                ucgml=liq.ucgml.copy(), ucgp=dyn["ucgp"].copy(), ucgforce=dyn["ucgforce"].copy())


COMPUTES = {"p": ("all", 1, ["ucgforce", "ucgvl", "ucgml"]), "s": ("half", GROUP_BIT, ["ucgl"]),
            "w": ("half", GROUP_BIT, ["ucgstate", "ucgp"])}

# name -> (group, columns, dump_modify lines (without "dump_modify ID"), timestep)
DUMP_CASES = {
    "ucg_plain": ("all", "id type x y z ucgstate ucgl ucgp", [], 0),
    "all_columns_sorted": ("all", "id mol type mass x y z xs ys zs vx vy vz fx fy fz ucgstate ucgl ucgp c_p[1] c_p[2] c_p[3] c_s c_w[1] c_w[2]",
                           ["sort id"], 250),
    "group_thresh": ("half", "id x ucgl ucgstate fx", ["thresh ucgl > 0.25 thresh x <= 3.5 thresh ucgstate == 1", "sort id"], 10),
    "thresh_ops": ("all", "id type mol vz", ["thresh type != 2 thresh mol >= 4 thresh vz < 0.3 thresh id |^ 0", "sort id"], 20),
    "nothing_selected": ("all", "id x", ["thresh x < -100.0"], 30),
    "user_formats": ("all", "id type x y fz ucgl ucgstate", ["format line \"%d %d %20.15g %g %g %g %d\" format float %14.8e format 4 %10.4f format int %6d",
                                                             "time yes units yes sort id"], 40),
}

# name -> (source dump columns, source modify, snapshots written (timesteps), read_dump words after the file name)
READ_CASES = {
    "replace_all": ("id type x y z vx vy vz fx fy fz ucgstate ucgl ucgp", ["sort id"], [0], "0 x y z vx vy vz fx fy fz ucgstate ucgl ucgp"),
    "scaled_second_snapshot": ("id xs ys zs ucgl ucgstate", [], [0, 10, 20], "10 x y z ucgl ucgstate box yes"),
    "ucg_only_box_no": ("id ucgp ucgl", [], [5], "5 ucgl ucgp box no"),
    "trim_subset": ("id x y z ucgl", ["thresh ucgl > 0.4"], [0], "0 x y z ucgl trim yes"),
    "trim_no_replace": ("id x y z ucgl", ["thresh x > 1.0"], [0], "0 x y z ucgl trim yes replace no"),
}


def second_state(synth):
    """the state read_dump is applied to: other coordinates/λ, a slightly different box, a few atoms missing"""
    liq, dyn = make_state(synth, seed=11)
    rng = np.random.default_rng(23)
    keep = np.sort(rng.permutation(liq.n)[: liq.n - 9])
    for k in ("x", "v", "type", "mask", "tag", "molecule", "ucgstate", "ucgl", "ucgvl", "ucgml"):
        setattr(liq, k, np.ascontiguousarray(getattr(liq, k)[keep]))
    for k in dyn:
        dyn[k] = np.ascontiguousarray(dyn[k][keep])
    liq.n = len(keep)
    liq.box_lo = liq.box_lo - 0.25
    liq.box_hi = liq.box_hi + 0.5
    liq.ucgl = rng.uniform(0, 1, liq.n)
    liq.x = rng.uniform(liq.box_lo, liq.box_hi, (liq.n, 3))
    return liq, dyn
