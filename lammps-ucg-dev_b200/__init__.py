"""ucg-b200: B200 (sm_100a) implementation of the LAMMPS UCG package's per-timestep hot path.

This Python module is a thin ctypes binding over the C-ABI in ``include/ucgb200.h``
(``libucgb200.so``, built in-tree by ``__graft_entry__.build()``).  It exists for the
tests and for ``bench.py``; the production consumers are the C++ LAMMPS style classes in
``host/``.  There is NO CPU fallback: importing works without a GPU (so that the symbol
table can be checked), but creating a :class:`Context` needs a CUDA device and a missing
library raises immediately.

The directory name contains hyphens, so import it through ``__graft_entry__.load_package()``
(alias ``lammps_ucg_dev_b200``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libucgb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ucgb200.h")

TAB_LOOKUP, TAB_LINEAR, TAB_SPLINE, TAB_BITMAP = 0, 1, 2, 3

F_X, F_V, F_F, F_TYPE, F_MASK, F_TAG, F_MOLECULE = (1 << k for k in range(7))
F_UCGSTATE, F_UCGL, F_UCGVL, F_UCGML, F_UCGP, F_UCGFORCE, F_SCORES, F_NUMSTATES = (1 << k for k in range(7, 15))
F_ALL = 0x7FFF

ERR_TEXT = {
    1: "Pair distance < table inner cutoff",
    2: "Pair distance > table outer cutoff",
    3: "neighbor row overflow",
    4: "density CV is defined for actual type 1 only",
    5: "atoms lost",
    6: "periodic box shorter than cut+skin",
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Atoms(C.Structure):
    _fields_ = [
        ("x", _dp), ("v", _dp), ("f", _dp),
        ("type", _ip), ("mask", _ip), ("tag", _ip), ("molecule", _ip),
        ("ucgstate", _ip),
        ("ucgl", _dp), ("ucgvl", _dp), ("ucgml", _dp), ("ucgp", _dp), ("ucgforce", _dp),
        ("ucgsoftmaxscores", _dp),
        ("num_ucgstates", _ip),
    ]


class Deck(C.Structure):
    _fields_ = [
        ("pair_style", C.c_int), ("nve", C.c_int), ("nve_groupbit", C.c_int), ("wall_bias", C.c_int),
        ("wall_barrier", C.c_double), ("langevin", C.c_int),
        ("t_start", C.c_double), ("t_stop", C.c_double), ("t_period", C.c_double),
        ("langevin_seed", C.c_int), ("langevin_groupbit", C.c_int),
        ("ucgstate", C.c_int), ("ucgstate_seed", C.c_int), ("ucgstate_rate", C.c_double),
        ("bethe_method", C.c_int), ("bethe_pseudo", C.c_int), ("bethe_prior", C.c_int),
        ("thermo_every", C.c_int), ("cluster_freq", C.c_int), ("post_force_order", C.c_int),
        ("bethe_noise_level", C.c_double), ("bethe_seed", C.c_int), ("langevin_bias", C.c_int), ("reserved", C.c_int * 2),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libucgb200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback for the UCG hot path.")
        _lib = C.CDLL(LIB_PATH)
        _lib.ucgb200_last_error.restype = C.c_char_p
        _lib.ucgb200_launch_count.restype = C.c_longlong
    return _lib


def declared_symbols() -> list[str]:
    """Every entry point declared in include/ucgb200.h."""
    import re
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"\b(ucgb200_[a-z_0-9]+)\s*\(", text)))


class UCGError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"ucgb200 rc={rc}: {msg}")
        self.rc = rc


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pi(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_ip) if a is not None else None


class Context:
    """One GPU context (one per process / per GPU)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._l = lib()
        self._h = C.c_void_p()
        rc = self._l.ucgb200_create(int(device), C.byref(self._h))
        if rc == -3:
            raise UCGError(rc, "no CUDA device: the UCG hot path has no CPU fallback")
        if rc:
            raise UCGError(rc, "ucgb200_create failed")
        if stream is not None:
            self._ck(self._l.ucgb200_set_stream(self._h, C.c_void_p(stream)))
        self.n_formal = 0
        self.n_actual = 0

    # ------------------------------------------------------------------ utils
    def _ck(self, rc: int) -> int:
        if rc < 0:
            raise UCGError(rc, self._l.ucgb200_last_error(self._h).decode())
        if rc > 0:
            msg = self._l.ucgb200_last_error(self._h).decode()
            raise UCGError(rc, ERR_TEXT.get(rc, "runtime condition") + (": " + msg if msg else ""))
        return rc

    def close(self):
        if self._h:
            self._l.ucgb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self._l.ucgb200_sync(self._h))

    def launch_count(self) -> int:
        return int(self._l.ucgb200_launch_count(self._h))

    def set_stream(self, stream: Optional[int]):
        self._ck(self._l.ucgb200_set_stream(self._h, C.c_void_p(stream or 0)))

    # ----------------------------------------------------------------- set-up
    def set_units(self, boltz=1.0, ftm2v=1.0, mvv2e=1.0):
        self._ck(self._l.ucgb200_set_units(self._h, C.c_double(boltz), C.c_double(ftm2v), C.c_double(mvv2e)))

    def set_box(self, lo, hi, periodic=(1, 1, 1)):
        lo, hi, per = _d(lo), _d(hi), _i(periodic)
        self._ck(self._l.ucgb200_set_box(self._h, _pd(lo), _pd(hi), _pi(per)))

    def set_subdomain(self, lo, hi):
        lo, hi = _d(lo), _d(hi)
        self._ck(self._l.ucgb200_set_subdomain(self._h, _pd(lo), _pd(hi)))

    def set_timestep(self, dt: float):
        self._ck(self._l.ucgb200_set_timestep(self._h, C.c_double(dt)))

    def set_special_lj(self, s):
        s = _d(s)
        self._ck(self._l.ucgb200_set_special_lj(self._h, _pd(s)))

    def set_types(self, n_actual, n_formal, n_states, formal_from_actual, chem_pot, mass):
        """1-based arrays: n_states[n_actual+1], formal_from_actual[(n_actual+1),2],
        chem_pot[n_formal+1], mass[n_formal+1]."""
        ns, ff, mu, m = _i(n_states), _i(formal_from_actual).reshape(-1), _d(chem_pot), _d(mass)
        assert ns.size == n_actual + 1 and ff.size == 2 * (n_actual + 1)
        assert mu.size == n_formal + 1 and m.size == n_formal + 1
        self.n_actual, self.n_formal = n_actual, n_formal
        self._ck(self._l.ucgb200_set_types(self._h, n_actual, n_formal, _pi(ns), _pi(ff), _pd(mu), _pd(m)))

    def set_kT(self, kT: float):
        self._ck(self._l.ucgb200_set_kT(self._h, C.c_double(kT)))

    def tables_clear(self):
        self._ck(self._l.ucgb200_tables_clear(self._h))

    def table_upload(self, tabstyle, tablength, innersq, delta, invdelta, deltasq6, cut, e, f,
                     e2=None, f2=None, rsq=None, drsq=None, de=None, df=None, nmask=0, nshiftbits=0) -> int:
        e, f = _d(e), _d(f)
        opt = [None if a is None else _d(a) for a in (e2, f2, rsq, drsq, de, df)]
        idx = C.c_int(-1)
        self._ck(self._l.ucgb200_table_upload(
            self._h, int(tabstyle), int(tablength), int(e.size), C.c_double(innersq), C.c_double(delta),
            C.c_double(invdelta), C.c_double(deltasq6), C.c_double(cut), int(nmask), int(nshiftbits),
            _pd(e), _pd(f), *[_pd(a) for a in opt], C.byref(idx)))
        return idx.value

    def set_pair_maps(self, tabindex, cutsq):
        ti, cs = _i(tabindex).reshape(-1), _d(cutsq).reshape(-1)
        nt = self.n_formal + 1
        assert ti.size == nt * nt and cs.size == nt * nt
        self._ck(self._l.ucgb200_set_pair_maps(self._h, _pi(ti), _pd(cs)))

    # ------------------------------------------------------------------ atoms
    @staticmethod
    def _atoms_struct(arrs: dict) -> Atoms:
        a = Atoms()
        for name, _ in Atoms._fields_:
            v = arrs.get(name)
            if v is None:
                continue
            setattr(a, name, v.ctypes.data_as(_dp if v.dtype == np.float64 else _ip))
        return a

    def atoms_upload(self, nlocal: int, **fields):
        """fields: x,v,f (n,3) float64; type,mask,tag,molecule,ucgstate int32; ucgl,... float64."""
        arrs, mask = {}, 0
        bits = dict(x=F_X, v=F_V, f=F_F, type=F_TYPE, mask=F_MASK, tag=F_TAG, molecule=F_MOLECULE,
                    ucgstate=F_UCGSTATE, ucgl=F_UCGL, ucgvl=F_UCGVL, ucgml=F_UCGML, ucgp=F_UCGP,
                    ucgforce=F_UCGFORCE, ucgsoftmaxscores=F_SCORES)
        for k, v in fields.items():
            if v is None:
                continue
            isint = k in ("type", "mask", "tag", "molecule", "ucgstate")
            arrs[k] = _i(v) if isint else _d(v)
            mask |= bits[k]
        st = self._atoms_struct(arrs)
        self._ck(self._l.ucgb200_atoms_upload(self._h, int(nlocal), C.byref(st), C.c_uint(mask)))
        self._keep = arrs

    def atoms_download(self, fields: Sequence[str]) -> dict:
        n = self.natoms()[0]
        shapes = dict(x=(n, 3), v=(n, 3), f=(n, 3), ucgsoftmaxscores=(n, 2))
        bits = dict(x=F_X, v=F_V, f=F_F, type=F_TYPE, mask=F_MASK, tag=F_TAG, molecule=F_MOLECULE,
                    ucgstate=F_UCGSTATE, ucgl=F_UCGL, ucgvl=F_UCGVL, ucgml=F_UCGML, ucgp=F_UCGP,
                    ucgforce=F_UCGFORCE, ucgsoftmaxscores=F_SCORES, num_ucgstates=F_NUMSTATES)
        arrs, mask = {}, 0
        for k in fields:
            isint = k in ("type", "mask", "tag", "molecule", "ucgstate", "num_ucgstates")
            arrs[k] = np.zeros(shapes.get(k, (n,)), dtype=np.int32 if isint else np.float64)
            mask |= bits[k]
        st = self._atoms_struct(arrs)
        self._ck(self._l.ucgb200_atoms_download(self._h, int(n), C.byref(st), C.c_uint(mask)))
        return arrs

    def atoms_download_into(self, **arrays):
        """same, into caller-owned C-contiguous arrays (pinned host memory makes the copies true DMA)"""
        n = self.natoms()[0]
        bits = dict(x=F_X, v=F_V, f=F_F, type=F_TYPE, mask=F_MASK, tag=F_TAG, molecule=F_MOLECULE,
                    ucgstate=F_UCGSTATE, ucgl=F_UCGL, ucgvl=F_UCGVL, ucgml=F_UCGML, ucgp=F_UCGP,
                    ucgforce=F_UCGFORCE, ucgsoftmaxscores=F_SCORES, num_ucgstates=F_NUMSTATES)
        mask = 0
        for k, v in arrays.items():
            isint = k in ("type", "mask", "tag", "molecule", "ucgstate", "num_ucgstates")
            assert v.flags.c_contiguous and v.dtype == (np.int32 if isint else np.float64) and v.shape[0] >= n, k
            mask |= bits[k]
        st = self._atoms_struct(arrays)
        self._ck(self._l.ucgb200_atoms_download(self._h, int(n), C.byref(st), C.c_uint(mask)))

    _BITS = dict(x=F_X, v=F_V, f=F_F, type=F_TYPE, mask=F_MASK, tag=F_TAG, molecule=F_MOLECULE, ucgstate=F_UCGSTATE,
                 ucgl=F_UCGL, ucgvl=F_UCGVL, ucgml=F_UCGML, ucgp=F_UCGP, ucgforce=F_UCGFORCE, ucgsoftmaxscores=F_SCORES)

    def step_host(self, inputs: dict, outputs: dict):
        """ucgb200_step_host: one timestep with host arrays in (`inputs`) and out (`outputs`, caller-owned C-contiguous
        arrays, ideally pinned); the device->host copies overlap the kernels of the step."""
        n = self.natoms()[0]
        im = om = 0
        for k, v in inputs.items():
            isint = k in ("type", "mask", "tag", "molecule", "ucgstate")
            assert v.flags.c_contiguous and v.dtype == (np.int32 if isint else np.float64) and v.shape[0] >= n, k
            im |= self._BITS[k]
        for k, v in outputs.items():
            isint = k in ("ucgstate",)
            assert v.flags.c_contiguous and v.dtype == (np.int32 if isint else np.float64) and v.shape[0] >= n, k
            om |= self._BITS[k]
        si, so = self._atoms_struct(inputs), self._atoms_struct(outputs)
        self._ck(self._l.ucgb200_step_host(self._h, C.byref(si), C.c_uint(im), C.byref(so), C.c_uint(om)))

    def natoms(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self._l.ucgb200_natoms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def force_clear(self):
        self._ck(self._l.ucgb200_force_clear(self._h))

    # --------------------------------------------------------------- neighbor
    def neigh_configure(self, skin: float, cut_override: float = 0.0):
        self._ck(self._l.ucgb200_neigh_configure(self._h, C.c_double(skin), C.c_double(cut_override)))

    def neigh_decide(self) -> int:
        f = C.c_int()
        self._ck(self._l.ucgb200_neigh_decide(self._h, C.byref(f)))
        return f.value

    def neigh_build(self):
        self._ck(self._l.ucgb200_neigh_build(self._h))

    def ghosts_forward(self):
        self._ck(self._l.ucgb200_ghosts_forward(self._h))

    def neigh_stats(self):
        t, m, b = C.c_longlong(), C.c_int(), C.c_int()
        self._ck(self._l.ucgb200_neigh_stats(self._h, C.byref(t), C.byref(m), C.byref(b)))
        return t.value, m.value, b.value

    def neigh_download(self):
        """-> dict(tag_i[n], numneigh[n], offsets[n+1], neigh_tags[T], neigh_shift[T])"""
        n, total = C.c_int(), C.c_longlong(0)
        self._ck(self._l.ucgb200_neigh_download(self._h, C.byref(n), C.byref(total), None, None, None, None, None))
        tag_i = np.zeros(n.value, np.int32)
        nn = np.zeros(n.value, np.int32)
        off = np.zeros(n.value + 1, np.int64)
        nt = np.zeros(max(total.value, 1), np.int32)
        sh = np.zeros(max(total.value, 1), np.int32)
        self._ck(self._l.ucgb200_neigh_download(
            self._h, C.byref(n), C.byref(total), _pi(tag_i), _pi(nn),
            off.ctypes.data_as(C.POINTER(C.c_longlong)), _pi(nt), _pi(sh)))
        return dict(tag_i=tag_i, numneigh=nn, offsets=off, neigh_tags=nt[:total.value], neigh_shift=sh[:total.value])

    # ------------------------------------------------------------ pair styles
    def pair_ucgld(self, eflag=0, vflag=0):
        self._ck(self._l.ucgb200_pair_ucgld(self._h, int(eflag), int(vflag)))

    def pair_bethe(self, eflag=0, vflag=0, method=1, pseudo=0, prior=0, noise=0.0, seed=1):
        self._ck(self._l.ucgb200_pair_bethe(self._h, int(eflag), int(vflag), int(method), int(pseudo),
                                            int(prior), C.c_double(noise), int(seed)))

    def pair_rleucg_configure(self, ntypes, actual_from_state, n_actual, n_states, use_entropy, cv_threshold,
                              threshold_radius, chem_pot, tabindex, cutsq, mass, kT):
        a = [_i(actual_from_state), _i(n_states), _i(use_entropy), _d(cv_threshold), _d(threshold_radius),
             _d(chem_pot), _i(tabindex).reshape(-1), _d(cutsq).reshape(-1), _d(mass)]
        self._ck(self._l.ucgb200_pair_rleucg_configure(
            self._h, int(ntypes), _pi(a[0]), int(n_actual), _pi(a[1]), _pi(a[2]), _pd(a[3]), _pd(a[4]), _pd(a[5]),
            _pi(a[6]), _pd(a[7]), _pd(a[8]), C.c_double(kT)))

    def pair_rleucg_probabilities(self):
        n = self.natoms()[0]
        p, f = np.zeros(n), np.zeros(n)
        self._ck(self._l.ucgb200_pair_rleucg_probabilities(self._h, int(n), _pd(p), _pd(f)))
        return p, f

    def pair_rleucg(self, eflag=0, vflag=0):
        self._ck(self._l.ucgb200_pair_rleucg(self._h, int(eflag), int(vflag)))

    def pair_bethe_density_configure(self, use_density, use_entropy, cv_threshold, threshold_radius):
        """per-actual-type arrays, 1-based (index 0 unused): pair_table_ucg_bethe_density.cpp:827-880"""
        a = [_i(use_density), _i(use_entropy), _d(cv_threshold), _d(threshold_radius)]
        self._ck(self._l.ucgb200_pair_bethe_density_configure(self._h, int(len(a[0]) - 1), _pi(a[0]), _pi(a[1]),
                                                              _pd(a[2]), _pd(a[3])))

    def pair_bethe_density(self, eflag=0, vflag=0):
        self._ck(self._l.ucgb200_pair_bethe_density(self._h, int(eflag), int(vflag)))

    def pair_bethe_density_priors(self):
        n = self.natoms()[0]
        p, f = np.zeros(n), np.zeros(n)
        self._ck(self._l.ucgb200_pair_bethe_density_priors(self._h, int(n), _pd(p), _pd(f)))
        return p, f

    def pair_peratom(self, eatom=True, vatom=True):
        """per-atom energy [n] / virial [n,6] of the last pair_ucgld(eflag | 2, vflag | 4) call, host order"""
        n = self.natoms()[0]
        e = np.zeros(n) if eatom else None
        v = np.zeros((n, 6)) if vatom else None
        self._ck(self._l.ucgb200_pair_peratom(self._h, int(n), _pd(e), _pd(v)))
        return e, v

    def pair_energy_virial(self):
        e, v = C.c_double(), np.zeros(6)
        self._ck(self._l.ucgb200_pair_energy_virial(self._h, C.byref(e), _pd(v)))
        return e.value, v

    # ------------------------------------------------------------------ fixes
    def fix_nve_initial(self, dtv, dtf, groupbit=1, wall=0):
        self._ck(self._l.ucgb200_fix_nve_initial(self._h, C.c_double(dtv), C.c_double(dtf), int(groupbit), int(wall)))

    def fix_nve_final(self, dtf, groupbit=1, wall=0):
        self._ck(self._l.ucgb200_fix_nve_final(self._h, C.c_double(dtf), int(groupbit), int(wall)))

    def fix_wall_bias(self, barrier, groupbit=1):
        self._ck(self._l.ucgb200_fix_wall_bias(self._h, C.c_double(barrier), int(groupbit)))

    def fix_ucgstate(self, mode=0, seed=1, rate=0.01, step=0):
        self._ck(self._l.ucgb200_fix_ucgstate(self._h, int(mode), int(seed), C.c_double(rate), C.c_longlong(step)))

    def fix_langevin(self, gfactor1, gfactor2, tsqrt, seed, step, groupbit=1, zero_v_skip=0):
        g1, g2 = _d(gfactor1), _d(gfactor2)
        self._ck(self._l.ucgb200_fix_langevin(self._h, _pd(g1), _pd(g2), int(g1.size - 1), C.c_double(tsqrt),
                                              int(seed), C.c_longlong(step), int(groupbit), int(zero_v_skip)))

    def lambda_ke(self, groupbit=1):
        ke, n = C.c_double(), C.c_longlong()
        self._ck(self._l.ucgb200_lambda_ke(self._h, int(groupbit), C.byref(ke), C.byref(n)))
        return ke.value, n.value

    def kinetic_energy(self, groupbit=1):
        ke, n = C.c_double(), C.c_longlong()
        self._ck(self._l.ucgb200_kinetic_energy(self._h, int(groupbit), C.byref(ke), C.byref(n)))
        return ke.value, n.value

    # ------------------------------------------------------------ multi-GPU halo
    def halo_configure(self, rank, nranks, procgrid):
        g = _i(procgrid)
        self._ck(self._l.ucgb200_halo_configure(self._h, int(rank), int(nranks), _pi(g)))
        self._nranks = int(nranks)

    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: the 128-byte ncclUniqueId to broadcast to every rank"""
        buf = C.create_string_buffer(128)
        n = lib().ucgb200_comm_unique_id(buf, 128)
        if n < 0:
            raise UCGError(n, "ucgb200_comm_unique_id: NCCL (libnccl.so.2) is not available")
        return buf.raw[:128]

    def comm_init(self, unique_id: bytes):
        """attach an NCCL communicator: setup()/run() then drive the brick exchanges themselves"""
        self._ck(self._l.ucgb200_comm_init(self._h, C.c_char_p(unique_id), len(unique_id)))

    def comm_stats(self):
        b, r, s_ = C.c_longlong(), C.c_int(), C.c_int()
        self._ck(self._l.ucgb200_comm_stats(self._h, C.byref(b), C.byref(r), C.byref(s_)))
        pm, pu = C.c_int(), C.c_longlong()
        self._ck(self._l.ucgb200_comm_transport(self._h, C.byref(pm), C.byref(pu)))
        return dict(bytes_forward=b.value, rebuilds=r.value, send_records=s_.value, pushes=pu.value,
                    transport="peer-mapped stores over NVLink (CUDA IPC), flags in the same push" if pm.value
                    else "NCCL send/recv groups + all-reduce of the rebuild flag")

    def comm_destroy(self):
        self._ck(self._l.ucgb200_comm_destroy(self._h))

    @staticmethod
    def halo_record_bytes():
        a, b, c_ = C.c_int(), C.c_int(), C.c_int()
        lib().ucgb200_halo_record_bytes(C.byref(a), C.byref(b), C.byref(c_))
        return dict(border=a.value, forward=b.value, migrate=c_.value)

    def migrate_prepare(self):
        counts = np.zeros(self._nranks, np.int32)
        self._ck(self._l.ucgb200_migrate_prepare(self._h, _pi(counts)))
        return counts

    def migrate_pack(self, d_ptr):
        self._ck(self._l.ucgb200_migrate_pack(self._h, C.c_void_p(d_ptr)))

    def migrate_unpack(self, d_ptr, nrecv):
        self._ck(self._l.ucgb200_migrate_unpack(self._h, C.c_void_p(d_ptr), int(nrecv)))

    def neigh_build_local(self):
        self._ck(self._l.ucgb200_neigh_build_local(self._h))

    def neigh_build_finish(self):
        self._ck(self._l.ucgb200_neigh_build_finish(self._h))

    def halo_send_counts(self):
        counts = np.zeros(self._nranks, np.int32)
        self._ck(self._l.ucgb200_halo_send_counts(self._h, _pi(counts)))
        return counts

    def halo_pack_border(self, d_ptr):
        self._ck(self._l.ucgb200_halo_pack_border(self._h, C.c_void_p(d_ptr)))

    def halo_unpack_border(self, d_ptr, recv_counts):
        rc = _i(recv_counts)
        self._ck(self._l.ucgb200_halo_unpack_border(self._h, C.c_void_p(d_ptr), _pi(rc)))

    def halo_pack_forward(self, d_ptr):
        self._ck(self._l.ucgb200_halo_pack_forward(self._h, C.c_void_p(d_ptr)))

    def halo_unpack_forward(self, d_ptr):
        self._ck(self._l.ucgb200_halo_unpack_forward(self._h, C.c_void_p(d_ptr)))

    # ---------------------------------------------------------- cluster switch
    def cluster_configure(self, mol_seed, mol_offset, cutoff, seed, prob_on, type_on, type_off, contact_pairs,
                          ntypes, groupbit=1):
        """fix ID group cluster_switch mol_seed mol_offset cutoff seed rateFreq N rateFile f contactFile f
        (fix_cluster_switch.cpp:36-170); atoms must be on the device already"""
        on, off = _i(type_on), _i(type_off)
        cp = _i(contact_pairs).reshape(-1)
        assert on.size == off.size and cp.size % 2 == 0
        self._ck(self._l.ucgb200_cluster_configure(self._h, int(mol_seed), int(mol_offset), C.c_double(cutoff), int(seed),
                                                   C.c_double(prob_on), int(on.size), _pi(on), _pi(off),
                                                   int(cp.size // 2), _pi(cp), int(ntypes), int(groupbit)))

    def cluster_check(self) -> int:
        n = C.c_int()
        self._ck(self._l.ucgb200_cluster_check(self._h, C.byref(n)))
        return n.value

    def cluster_switch(self):
        a, s_ = C.c_int(), C.c_int()
        self._ck(self._l.ucgb200_cluster_switch(self._h, C.byref(a), C.byref(s_)))
        return a.value, s_.value

    def cluster_stats(self) -> np.ndarray:
        out = np.zeros(8)
        self._ck(self._l.ucgb200_cluster_stats(self._h, _pd(out)))
        return out

    def cluster_get(self) -> dict:
        mm = C.c_int()
        self._ck(self._l.ucgb200_cluster_get(self._h, 0, None, None, None, None, C.byref(mm)))
        n = mm.value + 1
        a = {k: np.zeros(n, np.int32) for k in ("mol_cluster", "mol_state", "mol_restrict", "mol_accept")}
        self._ck(self._l.ucgb200_cluster_get(self._h, n, _pi(a["mol_cluster"]), _pi(a["mol_state"]), _pi(a["mol_restrict"]),
                                             _pi(a["mol_accept"]), C.byref(mm)))
        return a

    def deck_configure(self, **kw):
        d = Deck()
        for k, v in kw.items():
            setattr(d, k, v)
        self._ck(self._l.ucgb200_deck_configure(self._h, C.byref(d)))

    def setup(self):
        self._ck(self._l.ucgb200_setup(self._h))

    def run(self, nsteps: int):
        self._ck(self._l.ucgb200_run(self._h, int(nsteps)))

    def thermo(self) -> np.ndarray:
        out = np.zeros(16)
        self._ck(self._l.ucgb200_thermo(self._h, _pd(out)))
        return out

    def status(self):
        code, ti, tj, rsq = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self._l.ucgb200_status(self._h, C.byref(code), C.byref(ti), C.byref(tj), C.byref(rsq))
        return code.value, ti.value, tj.value, rsq.value

    def status_peek(self):
        """the sticky error word without clearing it"""
        code, ti, tj, rsq = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self._l.ucgb200_status_peek(self._h, C.byref(code), C.byref(ti), C.byref(tj), C.byref(rsq))
        return code.value, ti.value, tj.value, rsq.value

    def timers(self, enable: int = -1):
        ms = np.zeros(4)
        ln = np.zeros(4, np.int64)
        self._ck(self._l.ucgb200_timers(self._h, int(enable), _pd(ms), ln.ctypes.data_as(C.POINTER(C.c_longlong))))
        return dict(zip(("pair", "neigh", "comm", "modify"), ms.tolist())), dict(zip(("pair", "neigh", "comm", "modify"), ln.tolist()))

    def last_pair_ms(self) -> float:
        ms = C.c_double()
        self._ck(self._l.ucgb200_last_pair_ms(self._h, C.byref(ms)))
        return ms.value
