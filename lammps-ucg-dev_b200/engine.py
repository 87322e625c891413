"""Deck-level helpers over the C-ABI: what `pair_style table_ucgld ...`, `pair_coeff ...`,
`neighbor`, `fix ...` lines of an input deck translate to.  Used by tests, bench.py and
smoke(); mirrors the call sequence of the C++ style classes in ``host/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import Context, UCGError, lib, TAB_LINEAR

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class HostTable:
    """ucgb200_table: a tabulated potential built on the host (compute_table)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_file(cls, path, keyword, cut, tabstyle, tablength):
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_table_from_file(path.encode(), keyword.encode(), C.c_double(cut), int(tabstyle),
                                                int(tablength), C.byref(h), err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, r, e, f, cut, tabstyle, tablength, rflag=0, rlo=0.0, rhi=0.0, fprime=None):
        e = np.ascontiguousarray(e, np.float64)
        f = np.ascontiguousarray(f, np.float64)
        r = None if r is None else np.ascontiguousarray(r, np.float64)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_table_from_arrays(
            int(e.size), int(rflag), C.c_double(rlo), C.c_double(rhi), int(fprime is not None),
            C.c_double(fprime[0] if fprime else 0.0), C.c_double(fprime[1] if fprime else 0.0),
            r.ctypes.data_as(_dp) if r is not None else None, e.ctypes.data_as(_dp), f.ctypes.data_as(_dp),
            C.c_double(cut), int(tabstyle), int(tablength), C.byref(h), err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())
        return cls(h)

    def info(self):
        p = np.zeros(8)
        n = C.c_int()
        lib().ucgb200_host_table_info(self._h, p.ctypes.data_as(_dp), C.byref(n))
        return dict(innersq=p[0], delta=p[1], invdelta=p[2], deltasq6=p[3], cut=p[4], nmask=int(p[5]),
                    nshiftbits=int(p[6]), match=int(p[7]), n=n.value)

    def array(self, which):
        names = dict(rsq=0, e=1, f=2, de=3, df=4, e2=5, f2=6, drsq=7)
        cap = max(self.info()["n"], 1) + 8
        out = np.zeros(cap)
        n = lib().ucgb200_host_table_array(self._h, names[which], out.ctypes.data_as(_dp), int(cap))
        if n < 0:
            raise UCGError(n, "table_array")
        return out[:n].copy()

    def single(self, rsq, factor_lj=1.0):
        phi, ff = C.c_double(), C.c_double()
        rc = lib().ucgb200_host_table_single(self._h, C.c_double(rsq), C.c_double(factor_lj), C.byref(phi), C.byref(ff))
        return rc, phi.value, ff.value

    def upload(self, ctx: Context) -> int:
        idx = C.c_int(-1)
        ctx._ck(lib().ucgb200_host_table_upload(ctx._h, self._h, C.byref(idx)))
        return idx.value

    def __del__(self):
        try:
            if self._h:
                lib().ucgb200_host_table_free(self._h)
                self._h = None
        except Exception:
            pass


class StateMap:
    """ucgb200_statemap: read_state_settings + coeff + init_one bookkeeping."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_file(cls, path):
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_statemap_from_file(path.encode(), C.byref(h), err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())
        return cls(h)

    @classmethod
    def create(cls, n_actual, n_formal, n_states, formal_from_actual, chem_pot):
        ns = np.ascontiguousarray(n_states, np.int32)
        ff = np.ascontiguousarray(formal_from_actual, np.int32).reshape(-1)
        mu = np.ascontiguousarray(chem_pot, np.float64)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_statemap_create(int(n_actual), int(n_formal), ns.ctypes.data_as(_ip),
                                                ff.ctypes.data_as(_ip), mu.ctypes.data_as(_dp), C.byref(h), err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())
        return cls(h)

    def sizes(self):
        a, b = C.c_int(), C.c_int()
        lib().ucgb200_host_statemap_sizes(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def coeff(self, ilo, ihi, jlo, jhi, ns_i, ns_j, tables, cuts):
        t = np.ascontiguousarray(tables, np.int32)
        c = np.ascontiguousarray(cuts, np.float64)
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_statemap_coeff(self._h, ilo, ihi, jlo, jhi, ns_i, ns_j, t.ctypes.data_as(_ip),
                                               c.ctypes.data_as(_dp), err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())

    def init(self):
        err = C.create_string_buffer(512)
        rc = lib().ucgb200_host_statemap_init(self._h, err, 512)
        if rc:
            raise UCGError(rc, err.value.decode())

    def get(self):
        na, nf = self.sizes()
        ns = np.zeros(na + 1, np.int32)
        ff = np.zeros(2 * (na + 1), np.int32)
        mu = np.zeros(nf + 1)
        ti = np.zeros((nf + 1) ** 2, np.int32)
        cs = np.zeros((nf + 1) ** 2)
        lib().ucgb200_host_statemap_get(self._h, ns.ctypes.data_as(_ip), ff.ctypes.data_as(_ip), mu.ctypes.data_as(_dp),
                                        ti.ctypes.data_as(_ip), cs.ctypes.data_as(_dp))
        return dict(n_states=ns, formal_from_actual=ff.reshape(-1, 2), chem_pot=mu,
                    tabindex=ti.reshape(nf + 1, nf + 1), cutsq=cs.reshape(nf + 1, nf + 1))

    def apply(self, ctx: Context, mass):
        m = np.ascontiguousarray(mass, np.float64)
        na, nf = self.sizes()
        ctx.n_actual, ctx.n_formal = na, nf
        ctx._ck(lib().ucgb200_host_statemap_apply(ctx._h, self._h, m.ctypes.data_as(_dp)))

    def __del__(self):
        try:
            if self._h:
                lib().ucgb200_host_statemap_free(self._h)
                self._h = None
        except Exception:
            pass


def setup_single_type(ctx: Context, table_file, state_file, tabstyle=TAB_LINEAR, tablength=4096, cut=2.5,
                      skin=0.3, mass=1.0, dt=0.002, kT=1.0, box=None):
    """The deck
        units lj ; atom_style ucg ; pair_style table_ucgld <style> <N> <state_file>
        pair_coeff 1 1 2 2 <file> UCG_00 <cut> <file> UCG_01 <cut> <file> UCG_01 <cut> <file> UCG_11 <cut>
        neighbor <skin> bin ; timestep <dt>
    """
    ctx.set_units(1.0, 1.0, 1.0)
    if box is not None:
        ctx.set_box(box[0], box[1], (1, 1, 1))
    ctx.set_timestep(dt)
    sm = StateMap.from_file(state_file)
    idx, cuts = [], []
    for key in ("UCG_00", "UCG_01", "UCG_01", "UCG_11"):
        t = HostTable.from_file(table_file, key, cut, tabstyle, tablength)
        idx.append(t.upload(ctx))
        cuts.append(cut)
    sm.coeff(1, 1, 1, 1, 2, 2, idx, cuts)
    sm.init()
    na, nf = sm.sizes()
    sm.apply(ctx, np.array([0.0] + [mass] * nf))
    ctx.set_kT(kT)
    ctx.neigh_configure(skin)
    return sm


def upload_liquid(ctx: Context, liq):
    ctx.atoms_upload(liq.n, x=liq.x, v=liq.v, type=liq.type, mask=liq.mask, tag=liq.tag, molecule=liq.molecule,
                     ucgstate=liq.ucgstate, ucgl=liq.ucgl, ucgvl=liq.ucgvl, ucgml=liq.ucgml,
                     ucgp=np.full(liq.n, -1.0))
