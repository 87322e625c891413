"""ctypes binding of oracle/_ref/libucg_ref.so: the reference's own UCG/*.cpp, compiled verbatim
against the LAMMPS-API shim and driven by oracle/ref_driver.cpp.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libucg_ref.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
HOSTDRV = os.path.join(ROOT, "oracle", "_hostdrv", "libucg_hostdrv.so")   # same driver, product's GPU style classes
_libs = {}


def available() -> bool:
    return os.path.exists(LIB)


def hostdrv_available() -> bool:
    return os.path.exists(HOSTDRV)


def lib(path=None):
    path = path or LIB
    if path not in _libs:
        l = C.CDLL(path)
        l.ref_create.restype = C.c_void_p
        l.ref_error.restype = C.c_char_p
        l.ref_eng_vdwl.restype = C.c_double
        l.ref_fix_scalar.restype = C.c_double
        l.ref_fix_vector.restype = C.c_double
        l.ref_ntimestep.restype = C.c_longlong
        l.ref_neigh_pairs.restype = C.c_longlong
        _libs[path] = l
    return _libs[path]


def _d(a):
    return None if a is None else np.ascontiguousarray(a, np.float64)


def _i(a):
    return None if a is None else np.ascontiguousarray(a, np.int32)


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class RefSim:
    """A LAMMPS-like session running the reference's classes."""

    LIBPATH = None   # subclass HostSim points this at the product's host-class harness

    def __init__(self):
        self.l = lib(self.LIBPATH)
        self.h = C.c_void_p(self.l.ref_create())

    def __del__(self):
        try:
            if self.h:
                self.l.ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            msg = self.l.ref_error(self.h).decode()
            self.l.ref_clear_error(self.h)
            raise RuntimeError("reference: " + msg)

    def box(self, lo, hi, ntypes):
        lo, hi = _d(lo), _d(hi)
        self._ck(self.l.ref_box(self.h, _pd(lo), _pd(hi), int(ntypes)))

    def atoms(self, liq, ucgp=None):
        a = [_d(liq.x), _d(liq.v), _i(liq.type), _i(liq.mask), _i(liq.tag), _i(liq.molecule), _i(liq.ucgstate),
             _d(liq.ucgl), _d(liq.ucgvl), _d(liq.ucgml), _d(ucgp)]
        self._ck(self.l.ref_atoms(self.h, int(liq.n), _pd(a[0]), _pd(a[1]), _pi(a[2]), _pi(a[3]), _pi(a[4]), _pi(a[5]),
                                  _pi(a[6]), _pd(a[7]), _pd(a[8]), _pd(a[9]), _pd(a[10])))

    def set_state(self, x=None, v=None, ucgstate=None, ucgl=None, ucgvl=None, ucgp=None, f=None, ucgforce=None, scores=None):
        a = [_d(x), _d(v), _i(ucgstate), _d(ucgl), _d(ucgvl), _d(ucgp), _d(f), _d(ucgforce), _d(scores)]
        self._ck(self.l.ref_set_state(self.h, _pd(a[0]), _pd(a[1]), _pi(a[2]), _pd(a[3]), _pd(a[4]), _pd(a[5]), _pd(a[6]),
                                      _pd(a[7]), _pd(a[8])))

    def command(self, line: str):
        self._ck(self.l.ref_command(self.h, line.encode()))

    def nlocal(self):
        return self.l.ref_nlocal(self.h)

    def nghost(self):
        return self.l.ref_nghost(self.h)

    def get_atoms(self):
        n = self.nlocal()
        out = dict(x=np.zeros((n, 3)), v=np.zeros((n, 3)), f=np.zeros((n, 3)), type=np.zeros(n, np.int32),
                   tag=np.zeros(n, np.int32), ucgstate=np.zeros(n, np.int32), ucgl=np.zeros(n), ucgvl=np.zeros(n),
                   ucgp=np.zeros(n), ucgforce=np.zeros(n), ucgsoftmaxscores=np.zeros((n, 2)),
                   num_ucgstates=np.zeros(n, np.int32))
        self.l.ref_get_atoms(self.h, _pd(out["x"]), _pd(out["v"]), _pd(out["f"]), _pi(out["type"]), _pi(out["tag"]),
                             _pi(out["ucgstate"]), _pd(out["ucgl"]), _pd(out["ucgvl"]), _pd(out["ucgp"]),
                             _pd(out["ucgforce"]), _pd(out["ucgsoftmaxscores"]), _pi(out["num_ucgstates"]))
        return out

    def init(self):
        self._ck(self.l.ref_init(self.h))

    def compute_once(self, ev=1):
        self._ck(self.l.ref_compute_once(self.h, int(ev)))

    def setup(self, ev=1):
        self._ck(self.l.ref_setup(self.h, int(ev)))

    def run(self, nsteps, thermo_every=0):
        self._ck(self.l.ref_run(self.h, int(nsteps), int(thermo_every)))

    def fix_call(self, ifix, what):
        names = dict(initial_integrate=0, post_force=1, final_integrate=2, end_of_step=3, setup=4, pre_exchange=5,
                     min_post_force=6, post_force_respa_inner=7, post_force_respa_outer=8)
        self._ck(self.l.ref_fix_call(self.h, int(ifix), names[what]))

    def pair_peratom(self):
        """(eatom[n], vatom[n,6]) of the last compute_once(3), ghost tallies folded onto the owners"""
        n = self.nlocal()
        e, v = np.zeros(n), np.zeros((n, 6))
        self._ck(self.l.ref_pair_peratom(self.h, _pd(e), _pd(v)))
        return e, v

    def fix_call_respa(self, ifix, what, ilevel, iloop=0):
        names = dict(initial_integrate_respa=0, final_integrate_respa=1, post_force_respa=2)
        self._ck(self.l.ref_fix_call_respa(self.h, int(ifix), names[what], int(ilevel), int(iloop)))

    def min_energy_force(self, ev=0):
        """one force evaluation the way [stock] Min::energy_force does it: ... pair->compute, then every fix's min_post_force"""
        self._ck(self.l.ref_min_energy_force(self.h, int(ev)))

    def fix_scalar(self, ifix):
        return self.l.ref_fix_scalar(self.h, int(ifix))

    def fix_vector(self, ifix, k):
        return self.l.ref_fix_vector(self.h, int(ifix), int(k))

    def eng_vdwl(self):
        return self.l.ref_eng_vdwl(self.h)

    def virial(self):
        """(as shipped, per-pair tally): the reference leaves the former at zero"""
        a, b = np.zeros(6), np.zeros(6)
        self.l.ref_virial(self.h, _pd(a), _pd(b))
        return a, b

    def ntimestep(self):
        return self.l.ref_ntimestep(self.h)

    def nbuilds(self):
        return self.l.ref_nbuilds(self.h)

    def timers(self):
        t = np.zeros(4)
        self.l.ref_timers(self.h, _pd(t))
        return dict(zip(("pair", "neigh", "comm", "modify"), t.tolist()))

    def neigh_pairs(self, ilist=0):
        n = self.l.ref_neigh_pairs(self.h, int(ilist), None, None, C.c_longlong(0))
        a, b = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        self.l.ref_neigh_pairs(self.h, int(ilist), _pi(a), _pi(b), C.c_longlong(n))
        return a[:n], b[:n]

    def pair_single(self, itype, jtype, rsq, factor_lj=1.0):
        phi, ff = C.c_double(), C.c_double()
        self._ck(self.l.ref_pair_single(self.h, int(itype), int(jtype), C.c_double(rsq), C.c_double(factor_lj),
                                        C.byref(phi), C.byref(ff)))
        return phi.value, ff.value

    # --- decks -------------------------------------------------------------------
    @classmethod
    def single_type(cls, liq, table_file, state_file, pair="table_ucgld", tabstyle="linear", tablength=4096, cut=2.5,
                    skin=0.3, dt=0.002, extra="", ntypes=2):
        s = cls()
        s.box(liq.box_lo, liq.box_hi, ntypes)
        s.atoms(liq)
        s.command(f"neighbor {skin} bin")
        s.command(f"timestep {dt}")
        s.command(f"pair_style {pair} {tabstyle} {tablength} {state_file} {extra}")
        t = table_file
        s.command(f"pair_coeff 1 1 2 2 {t} UCG_00 {cut} {t} UCG_01 {cut} {t} UCG_01 {cut} {t} UCG_11 {cut}")
        return s

    @classmethod
    def ucgld_langevin(cls, liq, table_file, state_file, tablength=4096, dt=0.002, skin=0.3, t_start=1.0, t_stop=1.0,
                       t_period=1.0, seed=48291):
        """bench deck: nve/ucgld + ucgld/langevin + ucgstate ld"""
        s = cls.single_type(liq, table_file, state_file, tablength=tablength, dt=dt, skin=skin)
        s.command("fix 1 all nve/ucgld")
        s.command(f"fix 2 all ucgld/langevin {t_start} {t_stop} {t_period} {seed}")
        s.command("fix 3 all ucgstate ld")
        return s


class HostSim(RefSim):
    """The same LAMMPS-like session, but the styles are the product's GPU-backed classes
    (lammps-ucg-dev_b200/host/styles) calling libucgb200.so through the C-ABI."""
    LIBPATH = HOSTDRV
