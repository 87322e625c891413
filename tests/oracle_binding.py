"""ctypes binding of the CPU oracle (oracle/libucg_oracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libucg_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
        l = C.CDLL(LIB)
        l.orc_create.restype = C.c_void_p
        l.orc_error.restype = C.c_char_p
        l.orc_eng_vdwl.restype = C.c_double
        l.orc_lambda_temp.restype = C.c_double
        l.orc_neigh_total.restype = C.c_longlong
        l.orc_neigh_pairs.restype = C.c_longlong
        l.orc_ntimestep.restype = C.c_longlong
        for name in dir(l):
            pass
        _lib = l
    return _lib


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class Oracle:
    def __init__(self):
        self.l = lib()
        self.h = C.c_void_p(self.l.orc_create())
        self.n_formal = 0

    def __del__(self):
        try:
            if self.h:
                self.l.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def error(self) -> str:
        return self.l.orc_error(self.h).decode()

    def check(self):
        e = self.error()
        if e:
            raise RuntimeError("oracle: " + e)

    # set-up -----------------------------------------------------------------
    def set_units(self, boltz=1.0, ftm2v=1.0, mvv2e=1.0):
        self.l.orc_set_units(self.h, C.c_double(boltz), C.c_double(ftm2v), C.c_double(mvv2e))

    def set_box(self, lo, hi):
        lo, hi = _d(lo), _d(hi)
        self.l.orc_set_box(self.h, _pd(lo), _pd(hi))

    def set_dt(self, dt):
        self.l.orc_set_dt(self.h, C.c_double(dt))

    def set_newton(self, flag):
        self.l.orc_set_newton(self.h, int(flag))

    def set_special_lj(self, s):
        s = _d(s)
        self.l.orc_set_special_lj(self.h, _pd(s))

    def pair_style(self, tabstyle, tablength):
        self.l.orc_pair_style(self.h, int(tabstyle), int(tablength))

    def set_types(self, n_actual, n_formal, n_states, formal_from_actual, chem_pot, mass):
        self.n_formal = n_formal
        self.n_actual = n_actual
        a, b, c_, d = _i(n_states), _i(formal_from_actual).reshape(-1), _d(chem_pot), _d(mass)
        self.l.orc_set_types(self.h, n_actual, n_formal, _pi(a), _pi(b), _pd(c_), _pd(d))

    def table_add_file(self, path, keyword, cut) -> int:
        idx = self.l.orc_table_add_file(self.h, path.encode(), keyword.encode(), C.c_double(cut))
        self.check()
        return idx

    def table_add_arrays(self, r, e, f, cut, rflag=0, rlo=0.0, rhi=0.0, fprime=None) -> int:
        e, f = _d(e), _d(f)
        r = _d(r) if r is not None else None
        idx = self.l.orc_table_add_arrays(self.h, int(e.size), int(rflag), C.c_double(rlo), C.c_double(rhi),
                                          int(fprime is not None), C.c_double(fprime[0] if fprime else 0.0),
                                          C.c_double(fprime[1] if fprime else 0.0), _pd(r), _pd(e), _pd(f),
                                          C.c_double(cut))
        self.check()
        return idx

    def table_get(self, idx, which):
        names = dict(rsq=0, e=1, f=2, de=3, df=4, e2=5, f2=6, drsq=7)
        n = self.l.orc_table_len(self.h, idx)
        out = np.zeros(n + 8)
        m = self.l.orc_table_get(self.h, idx, names[which], _pd(out))
        return out[:m].copy()

    def table_params(self, idx):
        p = np.zeros(8)
        self.l.orc_table_params(self.h, idx, _pd(p))
        return dict(innersq=p[0], delta=p[1], invdelta=p[2], deltasq6=p[3], cut=p[4], nmask=int(p[5]),
                    nshiftbits=int(p[6]), match=int(p[7]))

    def pair_coeff(self, ilo, ihi, jlo, jhi, ns_i, ns_j, tables):
        t = _i(tables)
        self.l.orc_pair_coeff(self.h, ilo, ihi, jlo, jhi, ns_i, ns_j, _pi(t))
        self.check()

    def pair_init(self):
        self.l.orc_pair_init(self.h)
        self.check()

    def get_pair_maps(self):
        nt = self.n_formal + 1
        ti, cs = np.zeros(nt * nt, np.int32), np.zeros(nt * nt)
        self.l.orc_get_pair_maps(self.h, _pi(ti), _pd(cs))
        return ti.reshape(nt, nt), cs.reshape(nt, nt)

    def set_kT(self, kT):
        self.l.orc_set_kT(self.h, C.c_double(kT))

    # atoms ------------------------------------------------------------------
    def set_atoms(self, x, v, type, mask=None, tag=None, molecule=None, ucgstate=None, ucgl=None, ucgvl=None,
                  ucgml=None, ucgp=None):
        x = _d(x)
        n = x.shape[0]
        opt = lambda a, f: None if a is None else f(a)
        arrs = [x, opt(v, _d), _i(type), opt(mask, _i), opt(tag, _i), opt(molecule, _i), opt(ucgstate, _i),
                opt(ucgl, _d), opt(ucgvl, _d), opt(ucgml, _d), opt(ucgp, _d)]
        self.l.orc_set_atoms(self.h, n, _pd(arrs[0]), _pd(arrs[1]), _pi(arrs[2]), _pi(arrs[3]), _pi(arrs[4]),
                             _pi(arrs[5]), _pi(arrs[6]), _pd(arrs[7]), _pd(arrs[8]), _pd(arrs[9]), _pd(arrs[10]))

    def nlocal(self):
        return self.l.orc_nlocal(self.h)

    def nghost(self):
        return self.l.orc_nghost(self.h)

    def get_atoms(self):
        n = self.nlocal()
        out = dict(x=np.zeros((n, 3)), v=np.zeros((n, 3)), f=np.zeros((n, 3)), type=np.zeros(n, np.int32),
                   tag=np.zeros(n, np.int32), ucgstate=np.zeros(n, np.int32), ucgl=np.zeros(n), ucgvl=np.zeros(n),
                   ucgp=np.zeros(n), ucgforce=np.zeros(n), ucgsoftmaxscores=np.zeros((n, 2)),
                   num_ucgstates=np.zeros(n, np.int32))
        self.l.orc_get_atoms(self.h, _pd(out["x"]), _pd(out["v"]), _pd(out["f"]), _pi(out["type"]), _pi(out["tag"]),
                             _pi(out["ucgstate"]), _pd(out["ucgl"]), _pd(out["ucgvl"]), _pd(out["ucgp"]),
                             _pd(out["ucgforce"]), _pd(out["ucgsoftmaxscores"]), _pi(out["num_ucgstates"]))
        return out

    def set_forces(self, f=None, ucgforce=None, scores=None):
        f = None if f is None else _d(f)
        u = None if ucgforce is None else _d(ucgforce)
        s = None if scores is None else _d(scores)
        self.l.orc_set_forces(self.h, _pd(f), _pd(u), _pd(s))

    # neighbor / comm ---------------------------------------------------------
    def neigh_config(self, skin, full=0):
        self.l.orc_neigh_config(self.h, C.c_double(skin), int(full))

    def pbc(self):
        self.l.orc_pbc(self.h)

    def borders(self):
        self.l.orc_borders(self.h)

    def neigh_build(self):
        self.l.orc_neigh_build(self.h)

    def neigh_build_all(self):
        """pbc + borders + build, as on a Verlet rebuild step"""
        self.pbc()
        self.borders()
        self.neigh_build()

    def neigh_decide(self) -> int:
        return self.l.orc_neigh_decide(self.h)

    def forward_comm(self):
        self.l.orc_forward_comm(self.h)

    def reverse_comm(self):
        self.l.orc_reverse_comm(self.h)

    def force_clear(self):
        self.l.orc_force_clear(self.h)

    def neigh_pairs(self):
        t = self.l.orc_neigh_total(self.h)
        a, b = np.zeros(max(t, 1), np.int32), np.zeros(max(t, 1), np.int32)
        n = self.l.orc_neigh_pairs(self.h, _pi(a), _pi(b))
        return a[:n], b[:n]

    # pair --------------------------------------------------------------------
    def pair_ucgld(self, eflag=0, vflag=0):
        self.l.orc_pair_ucgld(self.h, int(eflag), int(vflag))
        self.check()

    def pair_bethe_config(self, method=1, pseudo=0, prior=0, noise=0.0, seed=1):
        self.l.orc_pair_bethe_config(self.h, int(method), int(pseudo), int(prior), C.c_double(noise), int(seed))

    def pair_bethe(self, eflag=0, vflag=0):
        self.l.orc_pair_bethe(self.h, int(eflag), int(vflag))
        self.check()

    def eng_vdwl(self):
        return self.l.orc_eng_vdwl(self.h)

    def virial(self):
        v = np.zeros(6)
        self.l.orc_virial(self.h, _pd(v))
        return v

    # fixes -------------------------------------------------------------------
    def fix_clear(self):
        self.l.orc_fix_clear(self.h)

    def fix_ttarget(self, T):
        self.l.orc_fix_ttarget(self.h, C.c_double(T))

    def fix_nve(self, groupbit=1):
        self.l.orc_fix_nve(self.h, int(groupbit))

    def fix_nve_wall(self, groupbit=1, bias_flag=0, barrier=0.1):
        self.l.orc_fix_nve_wall(self.h, int(groupbit), int(bias_flag), C.c_double(barrier))

    def fix_langevin(self, t_start, t_stop, t_period, seed, groupbit=1):
        self.l.orc_fix_langevin(self.h, int(groupbit), C.c_double(t_start), C.c_double(t_stop), C.c_double(t_period), int(seed))

    def fix_ucgstate(self, mode=0, seed=1, rate=0.01):
        self.l.orc_fix_ucgstate(self.h, int(mode), int(seed), C.c_double(rate))

    def lambda_temp(self):
        return self.l.orc_lambda_temp(self.h)

    def nve_initial(self, groupbit=1, wall=0):
        self.l.orc_nve_initial(self.h, int(groupbit), int(wall))

    def nve_final(self, groupbit=1, wall=0):
        self.l.orc_nve_final(self.h, int(groupbit), int(wall))

    def wall_bias(self, barrier, groupbit=1):
        self.l.orc_wall_bias(self.h, int(groupbit), C.c_double(barrier))

    def ucgstate_post_force(self, mode=0, rate=0.01):
        self.l.orc_ucgstate_post_force(self.h, int(mode), C.c_double(rate))

    # verlet ------------------------------------------------------------------
    def setup(self, eflag=1, vflag=1):
        self.l.orc_setup(self.h, int(eflag), int(vflag))
        self.check()

    def run(self, nsteps, thermo_every=0):
        self.l.orc_run(self.h, int(nsteps), int(thermo_every))
        self.check()

    def ntimestep(self):
        return self.l.orc_ntimestep(self.h)

    def nbuilds(self):
        return self.l.orc_nbuilds(self.h)

    def timers(self):
        t = np.zeros(4)
        self.l.orc_timers(self.h, _pd(t))
        return dict(zip(("pair", "neigh", "comm", "modify"), t.tolist()))

    # convenience ---------------------------------------------------------------
    @classmethod
    def single_type(cls, liq, table_file, tabstyle=1, tablength=4096, cut=2.5, skin=0.3, mu=(0.0, 0.5),
                    mass=1.0, dt=0.002, kT=1.0, full=0):
        """Same deck as lammps_ucg_dev_b200.engine.setup_single_type + upload_liquid."""
        o = cls()
        o.set_units(1.0, 1.0, 1.0)
        o.set_box(liq.box_lo, liq.box_hi)
        o.set_dt(dt)
        o.pair_style(tabstyle, tablength)
        o.set_types(1, 2, [0, 2], [[0, 0], [1, 2]], [0.0, mu[0], mu[1]], [0.0, mass, mass])
        idx = [o.table_add_file(table_file, k, cut) for k in ("UCG_00", "UCG_01", "UCG_01", "UCG_11")]
        o.pair_coeff(1, 1, 1, 1, 2, 2, idx)
        o.pair_init()
        o.set_kT(kT)
        o.neigh_config(skin, full)
        o.set_atoms(liq.x, liq.v, liq.type, liq.mask, liq.tag, liq.molecule, liq.ucgstate, liq.ucgl, liq.ucgvl,
                    liq.ucgml, np.full(liq.n, -1.0))
        return o


def ranmars(seed, n):
    out = np.zeros(n)
    lib().orc_ranmars_fill(int(seed), int(n), _pd(out))
    return out


def ranpark(seed, n):
    out = np.zeros(n)
    lib().orc_ranpark_fill(int(seed), int(n), _pd(out))
    return out
