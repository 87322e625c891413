"""ctypes binding of the dump / read_dump / read_data taps (include/ucgb200.h "dump / read_dump taps",
include/ucgb200_host.h "dump custom / read_dump / read_data").  The command-level classes take the same
words a LAMMPS deck would carry after `dump`, `dump_modify`, `compute ... property/atom`, `read_dump`.

Reference: dump_custom.cpp, read_dump.cpp, reader_native.cpp (patched stock files of the reference tree),
UCG/atom_vec_ucg.cpp:85-90, 145-234.
"""
from __future__ import annotations

import ctypes as C
import shlex
from typing import Optional, Sequence

import numpy as np

from . import Atoms, Context, UCGError, lib

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

COL = dict(id=0, mol=1, type=2, mass=3, x=4, y=5, z=6, xs=7, ys=8, zs=9, vx=10, vy=11, vz=12, fx=13, fy=14, fz=15,
           ucgstate=16, ucgl=17, ucgp=18, proc=19, q=20,
           p_ucgstate=21, p_ucgl=22, p_ucgforce=23, p_ucgvl=24, p_ucgp=25, p_ucgml=26)
INT_COLS = {COL[k] for k in ("id", "mol", "type", "proc", "ucgstate")}
THRESH_OPS = {"<": 0, "<=": 1, ">": 2, ">=": 3, "==": 4, "!=": 5, "|^": 6}
ORDER_INDEX, ORDER_ID = 0, 1


class DumpSpec(C.Structure):
    _fields_ = [("ncols", C.c_int), ("cols", _ip), ("col_groupbit", _ip), ("groupbit", C.c_int), ("nthresh", C.c_int),
                ("thresh_col", _ip), ("thresh_op", _ip), ("thresh_value", _dp), ("order", C.c_int)]


def _spec(cols, groupbit=1, col_groupbit=None, thresh=(), order=ORDER_INDEX):
    """thresh: sequence of (column name or code, operator string, value)"""
    code = lambda c: COL[c] if isinstance(c, str) else int(c)
    keep = dict(cols=np.ascontiguousarray([code(c) for c in cols], np.int32))
    keep["bits"] = np.ascontiguousarray(col_groupbit if col_groupbit is not None else [-1] * len(cols), np.int32)
    keep["tc"] = np.ascontiguousarray([code(t[0]) for t in thresh] or [0], np.int32)
    keep["to"] = np.ascontiguousarray([THRESH_OPS[t[1]] for t in thresh] or [0], np.int32)
    keep["tv"] = np.ascontiguousarray([float(t[2]) for t in thresh] or [0.0], np.float64)
    sp = DumpSpec(len(cols), keep["cols"].ctypes.data_as(_ip), keep["bits"].ctypes.data_as(_ip), int(groupbit), len(thresh),
                  keep["tc"].ctypes.data_as(_ip), keep["to"].ctypes.data_as(_ip), keep["tv"].ctypes.data_as(_dp), int(order))
    return sp, keep


def dump_count(ctx: Context, cols, **kw) -> int:
    sp, keep = _spec(cols, **kw)
    n = C.c_longlong(0)
    ctx._ck(ctx._l.ucgb200_dump_count(ctx._h, C.byref(sp), C.byref(n)))
    return n.value


def dump_pack(ctx: Context, cols, **kw) -> np.ndarray:
    """DumpCustom::count + pack (+ sort): the (nchoose, ncols) double buffer"""
    sp, keep = _spec(cols, **kw)
    cap = ctx.natoms()[0]
    buf = np.zeros((max(cap, 1), len(cols)))
    n = C.c_longlong(0)
    ctx._ck(ctx._l.ucgb200_dump_pack(ctx._h, C.byref(sp), buf.ctypes.data_as(_dp), C.c_longlong(cap), C.byref(n)))
    return buf[:n.value]


def dump_text(ctx: Context, cols, **kw) -> bytes:
    """the same rows as text, formatted on the device with the default formats (%d / %g)"""
    sp, keep = _spec(cols, **kw)
    n, nb = C.c_longlong(0), C.c_longlong(0)
    ctx._ck(ctx._l.ucgb200_dump_text(ctx._h, C.byref(sp), None, C.c_longlong(0), C.byref(n), C.byref(nb)))
    out = C.create_string_buffer(max(nb.value, 1))
    ctx._ck(ctx._l.ucgb200_dump_text_copy(ctx._h, out, C.c_longlong(nb.value)))
    return out.raw[:nb.value]


def update_by_tag(ctx: Context, fieldtypes, fields, scaled=False, snap_lo=None, snap_hi=None):
    """ReadDump::process_atoms (replace mode): returns (updated flags in host order, nreplace)"""
    code = lambda c: COL[c] if isinstance(c, str) else int(c)
    ft = np.ascontiguousarray([code(c) for c in fieldtypes], np.int32)
    f = np.ascontiguousarray(fields, np.float64).reshape(-1, ft.size)
    lo = None if snap_lo is None else np.ascontiguousarray(snap_lo, np.float64)
    hi = None if snap_hi is None else np.ascontiguousarray(snap_hi, np.float64)
    upd = np.zeros(max(ctx.natoms()[0], 1), np.int32)
    nrep = C.c_longlong(0)
    ctx._ck(ctx._l.ucgb200_atoms_update_by_tag(
        ctx._h, int(f.shape[0]), int(ft.size), ft.ctypes.data_as(_ip), f.ctypes.data_as(_dp), int(bool(scaled)),
        None if lo is None else lo.ctypes.data_as(_dp), None if hi is None else hi.ctypes.data_as(_dp),
        upd.ctypes.data_as(_ip), C.byref(nrep)))
    return upd[:ctx.natoms()[0]], nrep.value


def _argv(words: Sequence[str]):
    arr = (C.c_char_p * len(words))(*[w.encode() for w in words])
    return arr


class DumpCustom:
    """`dump ID group custom N file col ...` on a device context; `modify()` takes dump_modify's words"""

    def __init__(self, ctx: Context, line: str, groupbit: int = 1):
        self.ctx, self._l = ctx, lib()
        words = shlex.split(line)
        if words and words[0] == "dump":
            words = words[1:]
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        if self._l.ucgb200_host_dump_create(len(words), _argv(words), int(groupbit), C.byref(self._h), err, 512):
            raise UCGError(-1, err.value.decode())

    def close(self):
        if self._h:
            self._l.ucgb200_host_dump_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, fn, *args):
        err = C.create_string_buffer(512)
        if fn(self._h, *args, err, 512):
            raise UCGError(-1, err.value.decode())

    def bind_compute(self, line: str, groupbit: int = 1):
        """`compute ID group property/atom name ...`"""
        w = shlex.split(line)
        if w and w[0] == "compute":
            w = w[1:]
        if len(w) < 4 or w[2] != "property/atom":
            raise UCGError(-1, "Illegal compute property/atom command")
        names = w[3:]
        self._call(self._l.ucgb200_host_dump_bind_compute, w[0].encode(), int(groupbit), len(names), _argv(names))

    def modify(self, line: str):
        w = shlex.split(line)   # a quoted format line is one word, as in a LAMMPS deck
        if w and w[0] == "dump_modify":
            w = w[2:]
        self._call(self._l.ucgb200_host_dump_modify, len(w), _argv(w))

    def write(self, ntimestep: int, time: float = 0.0, units: str = "lj"):
        self._call(self._l.ucgb200_host_dump_write, self.ctx._h, C.c_longlong(int(ntimestep)), C.c_double(time), units.encode())

    def stats(self):
        r, b, e = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
        self._l.ucgb200_host_dump_stats(self._h, C.byref(r), C.byref(b), C.byref(e))
        return dict(rows=r.value, bytes=b.value, nevery=e.value)


def run(ctx: Context, nsteps: int, dumps: Sequence[DumpCustom] = (), dt: float = 0.005, units: str = "lj"):
    """`run N` of a configured resident deck with its dumps ([stock] Output scheduling)"""
    arr = (C.c_void_p * max(len(dumps), 1))(*[d._h for d in dumps])
    err = C.create_string_buffer(512)
    if lib().ucgb200_host_run(ctx._h, C.c_longlong(int(nsteps)), len(dumps), arr, C.c_double(dt), units.encode(), err, 512):
        raise UCGError(-1, err.value.decode())


def read_dump(ctx: Context, line: str) -> dict:
    """`read_dump file Nstep field ... keyword value ...` applied to the resident atoms"""
    w = shlex.split(line)
    if w and w[0] == "read_dump":
        w = w[1:]
    stats = (C.c_longlong * 7)()
    err = C.create_string_buffer(512)
    if lib().ucgb200_host_read_dump(ctx._h, len(w), _argv(w), stats, err, 512):
        raise UCGError(-1, err.value.decode())
    keys = ("before", "snapshot", "purged", "replaced", "trimmed", "added", "after")
    return dict(zip(keys, [int(v) for v in stats]))


class DataFile:
    """a data file of atom_style ucg, parsed on the host (read_data + AtomVecUCG::data_atom_post)"""

    def __init__(self, path: str):
        self._l = lib()
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        if self._l.ucgb200_host_data_read(path.encode(), C.byref(self._h), err, 512):
            raise UCGError(-1, err.value.decode())
        n, nt = C.c_longlong(0), C.c_int(0)
        lo, hi = np.zeros(3), np.zeros(3)
        self._l.ucgb200_host_data_info(self._h, C.byref(n), C.byref(nt), lo.ctypes.data_as(_dp), hi.ctypes.data_as(_dp))
        self.natoms, self.ntypes, self.box_lo, self.box_hi = n.value, nt.value, lo, hi

    def arrays(self) -> dict:
        view = Atoms()
        q, img, mass = _dp(), _ip(), _dp()
        self._l.ucgb200_host_data_view(self._h, C.byref(view), C.byref(q), C.byref(img), C.byref(mass))
        n = self.natoms
        as_d = lambda p, m: np.ctypeslib.as_array(p, shape=(m,)).copy()
        out = dict(x=as_d(view.x, 3 * n).reshape(n, 3), v=as_d(view.v, 3 * n).reshape(n, 3), q=as_d(q, n),
                   ucgl=as_d(view.ucgl, n), ucgvl=as_d(view.ucgvl, n), ucgml=as_d(view.ucgml, n), ucgp=as_d(view.ucgp, n),
                   mass=as_d(mass, self.ntypes + 1))
        for k, p in (("tag", view.tag), ("molecule", view.molecule), ("type", view.type), ("ucgstate", view.ucgstate),
                     ("mask", view.mask), ("image", img)):
            out[k] = np.ctypeslib.as_array(p, shape=(n,)).copy()
        return out

    def upload(self, ctx: Context, periodic=(1, 1, 1)):
        per = np.ascontiguousarray(periodic, np.int32)
        ctx._ck(self._l.ucgb200_host_data_upload(ctx._h, self._h, per.ctypes.data_as(_ip)))

    def close(self):
        if self._h:
            self._l.ucgb200_host_data_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
