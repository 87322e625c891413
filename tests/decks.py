"""Deck builders shared by the GPU parity tests: the same deck set up on the GPU context
(through the product's host helpers) and on the oracle."""
import numpy as np

import oracle_binding as ob


def gpu_single_type(pkg, liq, fixtures, tabstyle=1, tablength=4096, table="table4096", cut=2.5, skin=0.3,
                    dt=0.002, kT=1.0):
    from lammps_ucg_dev_b200 import engine
    ctx = pkg.Context(0)
    engine.setup_single_type(ctx, fixtures[table], fixtures["state"], tabstyle=tabstyle, tablength=tablength,
                             cut=cut, skin=skin, dt=dt, kT=kT, box=(liq.box_lo, liq.box_hi))
    engine.upload_liquid(ctx, liq)
    return ctx


def orc_single_type(liq, fixtures, tabstyle=1, tablength=4096, table="table4096", cut=2.5, skin=0.3, dt=0.002,
                    kT=1.0, full=0):
    return ob.Oracle.single_type(liq, fixtures[table], tabstyle=tabstyle, tablength=tablength, cut=cut, skin=skin,
                                 dt=dt, kT=kT, full=full)


# mixed deck: actual type 1 = plain CG site (1 state, formal 1), actual type 2 = UCG site with
# formal types (2,3), chemical potentials (0, 0.3).  Exercises scenarios 1-4.
MIXED = dict(n_actual=2, n_formal=3, n_states=[0, 1, 2], formal=[[0, 0], [1, 0], [2, 3]], mu=[0.0, 0.0, 0.0, 0.3],
             mass=[0.0, 1.0, 2.0, 2.0])
# pair_coeff lines: (ilo,ihi,jlo,jhi,Ns_i,Ns_j,[table keywords])
MIXED_COEFF = [
    (1, 1, 1, 1, 1, 1, ["UCG_00"]),
    (1, 1, 2, 2, 1, 2, ["UCG_01", "UCG_11"]),
    (2, 2, 2, 2, 2, 2, ["UCG_00", "UCG_01", "UCG_01", "UCG_11"]),
]


def mixed_types(liq, frac_cg=0.3, seed=7):
    t = np.where(np.random.default_rng(seed).uniform(size=liq.n) < frac_cg, 1, 2).astype(np.int32)
    liq.type = t
    return liq


def gpu_mixed(pkg, liq, fixtures, tabstyle=1, tablength=1024, table="table1024", cut=2.5, skin=0.3, kT=1.0):
    from lammps_ucg_dev_b200 import engine
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    ctx.set_timestep(0.002)
    sm = engine.StateMap.create(MIXED["n_actual"], MIXED["n_formal"], MIXED["n_states"], MIXED["formal"], MIXED["mu"])
    for ilo, ihi, jlo, jhi, nsi, nsj, keys in MIXED_COEFF:
        idx = [engine.HostTable.from_file(fixtures[table], k, cut, tabstyle, tablength).upload(ctx) for k in keys]
        sm.coeff(ilo, ihi, jlo, jhi, nsi, nsj, idx, [cut] * len(idx))
    sm.init()
    sm.apply(ctx, MIXED["mass"])
    ctx.set_kT(kT)
    ctx.neigh_configure(skin)
    engine.upload_liquid(ctx, liq)
    return ctx


def orc_mixed(liq, fixtures, tabstyle=1, tablength=1024, table="table1024", cut=2.5, skin=0.3, kT=1.0):
    o = ob.Oracle()
    o.set_units(1.0, 1.0, 1.0)
    o.set_box(liq.box_lo, liq.box_hi)
    o.set_dt(0.002)
    o.pair_style(tabstyle, tablength)
    o.set_types(MIXED["n_actual"], MIXED["n_formal"], MIXED["n_states"], MIXED["formal"], MIXED["mu"], MIXED["mass"])
    for ilo, ihi, jlo, jhi, nsi, nsj, keys in MIXED_COEFF:
        idx = [o.table_add_file(fixtures[table], k, cut) for k in keys]
        o.pair_coeff(ilo, ihi, jlo, jhi, nsi, nsj, idx)
    o.pair_init()
    o.set_kT(kT)
    o.neigh_config(skin, 0)
    o.set_atoms(liq.x, liq.v, liq.type, liq.mask, liq.tag, liq.molecule, liq.ucgstate, liq.ucgl, liq.ucgvl, liq.ucgml,
                np.full(liq.n, -1.0))
    return o


def oracle_forces(o, eflag=1, vflag=1, pair="ucgld"):
    o.neigh_build_all()
    o.force_clear()
    getattr(o, "pair_" + pair)(eflag, vflag)
    o.reverse_comm()
    return o.get_atoms()


def rel_err(a, b):
    """max |a-b| / max |b| over the WHOLE array: how the north star's "within 1e-6 relative" is read throughout the
    suite (an error is measured against the largest force in the system, not against each site's own, which may be
    arbitrarily close to zero).  tests/test_gpu_config_sizes.py additionally asserts the worst PER-SITE ratio over the
    sites whose reference magnitude exceeds 1e-3 of the largest (per_site_rel)."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
