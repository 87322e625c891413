"""Synthetic UCG liquids, LAMMPS table files and state-settings files.

The reference ships no inputs at all, so the benchmark systems are defined here:
fcc lattice at rho* = 0.8442 with n^3 cells x 4 sites, jittered by U(-0.05,0.05) sigma,
one 2-state UCG type (formal types 1,2), LJ 12-6 tables shifted to 0 at rc = 2.5 with
(eps,sigma) = (1,1) for 0-0, (0.8,1.025) for 0-1 = 1-0 and (0.6,1.05) for 1-1, written as
``N <len> RSQ 0.5 2.5`` table sections.  Pure numpy; used by tests and bench.py.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

RHO = 0.8442
RC = 2.5
RLO = 0.5
LJ_PARAMS = {"00": (1.0, 1.0), "01": (0.8, 1.025), "11": (0.6, 1.05)}


@dataclass
class Liquid:
    n: int
    box_lo: np.ndarray
    box_hi: np.ndarray
    x: np.ndarray
    v: np.ndarray
    type: np.ndarray
    mask: np.ndarray
    tag: np.ndarray
    molecule: np.ndarray
    ucgstate: np.ndarray
    ucgl: np.ndarray
    ucgvl: np.ndarray
    ucgml: np.ndarray
    meta: dict = field(default_factory=dict)


def fcc_liquid(ncell, rho=RHO, jitter=0.05, T=1.0, ucgml=10.0, seed=12345, mol_size=0) -> Liquid:
    """n^3 fcc cells (ncell may be an int or a 3-tuple) -> 4*nx*ny*nz sites."""
    if np.isscalar(ncell):
        ncell = (int(ncell),) * 3
    nx, ny, nz = ncell
    a = (4.0 / rho) ** (1.0 / 3.0)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    x = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    n = x.shape[0]
    rng = np.random.default_rng(seed)
    x += rng.uniform(-jitter, jitter, size=x.shape)
    hi = np.array([nx, ny, nz], dtype=np.float64) * a
    x = np.mod(x, hi)
    x = np.minimum(x, np.nextafter(hi, 0.0))
    ucgl = np.random.default_rng(seed + 2).uniform(0.0, 1.0, n)
    ucgstate = (ucgl >= 0.5).astype(np.int32)
    ml = np.full(n, float(ucgml))
    ucgvl = np.random.default_rng(seed + 3).normal(0.0, np.sqrt(T / ucgml), n)
    v = np.random.default_rng(seed + 4).normal(0.0, np.sqrt(T), (n, 3))
    v -= v.mean(axis=0)
    tag = np.arange(1, n + 1, dtype=np.int32)
    mol = np.zeros(n, np.int32) if mol_size <= 0 else ((tag - 1) // mol_size + 1).astype(np.int32)
    return Liquid(n=n, box_lo=np.zeros(3), box_hi=hi, x=np.ascontiguousarray(x), v=np.ascontiguousarray(v),
                  type=np.ones(n, np.int32), mask=np.ones(n, np.int32), tag=tag, molecule=mol,
                  ucgstate=ucgstate, ucgl=ucgl, ucgvl=ucgvl, ucgml=ml,
                  meta=dict(ncell=ncell, rho=rho, a=a, T=T, seed=seed))


def _splitmix64(x):
    """vectorised SplitMix64: counter -> 64 well-mixed bits (decomposition-independent streams)"""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(ids, stream):
    with np.errstate(over="ignore"):
        bits = _splitmix64(ids.astype(np.uint64) * np.uint64(64) + np.uint64(stream))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _normal(ids, stream):
    u1 = np.maximum(_uniform(ids, stream), 1e-300)
    u2 = _uniform(ids, stream + 1)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def fcc_liquid_brick(ncell, grid, rank, rho=RHO, jitter=0.05, T=1.0, ucgml=10.0) -> Liquid:
    """The sites of brick `rank` of the global ncell fcc liquid (grid = bricks per dimension).
    Every per-site random number is a hash of the GLOBAL site id, so the global system does not
    depend on the decomposition; jitter may push a site across a brick face — the first
    rebuild's migration hands it to its owner."""
    nx, ny, nz = ncell
    gx, gy, gz = grid
    cx, cy, cz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
    a = (4.0 / rho) ** (1.0 / 3.0)
    rx = np.arange(cx * nx // gx, (cx + 1) * nx // gx)
    ry = np.arange(cy * ny // gy, (cy + 1) * ny // gy)
    rz = np.arange(cz * nz // gz, (cz + 1) * nz // gz)
    ii, jj, kk = np.meshgrid(rx, ry, rz, indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    cid = (cells[:, 0].astype(np.int64) * ny + cells[:, 1]) * nz + cells[:, 2]
    gid = (cid[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
    x = (cells[:, None, :].astype(np.float64) + basis[None, :, :]).reshape(-1, 3) * a
    for d in range(3):
        x[:, d] += (2.0 * _uniform(gid, 1 + d) - 1.0) * jitter
    hi = np.array([nx, ny, nz], dtype=np.float64) * a
    x = np.mod(x, hi)
    x = np.minimum(x, np.nextafter(hi, 0.0))
    n = x.shape[0]
    ucgl = _uniform(gid, 4)
    v = np.stack([_normal(gid, 10), _normal(gid, 12), _normal(gid, 14)], axis=1) * np.sqrt(T)
    ucgvl = _normal(gid, 16) * np.sqrt(T / ucgml)
    return Liquid(n=n, box_lo=np.zeros(3), box_hi=hi, x=np.ascontiguousarray(x), v=np.ascontiguousarray(v),
                  type=np.ones(n, np.int32), mask=np.ones(n, np.int32), tag=(gid + 1).astype(np.int32),
                  molecule=np.zeros(n, np.int32), ucgstate=(ucgl >= 0.5).astype(np.int32), ucgl=ucgl, ucgvl=ucgvl,
                  ucgml=np.full(n, float(ucgml)), meta=dict(ncell=ncell, grid=grid, rank=rank, a=a))


def lj_table(eps, sigma, npts, rlo=RLO, rhi=RC):
    """(r, e, f) on the RSQ grid of a LAMMPS `N npts RSQ rlo rhi` section; e shifted to 0 at rhi."""
    i = np.arange(npts, dtype=np.float64)
    rsq = rlo * rlo + (rhi * rhi - rlo * rlo) * i / (npts - 1)
    r = np.sqrt(rsq)
    sr6 = (sigma / r) ** 6
    src6 = (sigma / rhi) ** 6
    e = 4.0 * eps * (sr6 * sr6 - sr6) - 4.0 * eps * (src6 * src6 - src6)
    f = 24.0 * eps * (2.0 * sr6 * sr6 - sr6) / r
    return r, e, f


def write_table_file(path, npts=4096, rlo=RLO, rhi=RC, params=None, style="RSQ"):
    """One file with sections UCG_00, UCG_01, UCG_11."""
    params = params or LJ_PARAMS
    with open(path, "w") as fp:
        fp.write("# synthetic UCG tables (ucg-b200 fixtures)\n\n")
        for key, (eps, sig) in params.items():
            if style == "RSQ":
                r, e, f = lj_table(eps, sig, npts, rlo, rhi)
                fp.write(f"UCG_{key}\nN {npts} RSQ {rlo!r} {rhi!r}\n\n")
            else:
                r = np.linspace(rlo, rhi, npts)
                sr6 = (sig / r) ** 6
                src6 = (sig / rhi) ** 6
                e = 4.0 * eps * (sr6 * sr6 - sr6) - 4.0 * eps * (src6 * src6 - src6)
                f = 24.0 * eps * (2.0 * sr6 * sr6 - sr6) / r
                fp.write(f"UCG_{key}\nN {npts} R {rlo!r} {rhi!r}\n\n")
            for k in range(npts):
                fp.write("%d %s %s %s\n" % (k + 1, repr(float(r[k])), repr(float(e[k])), repr(float(f[k]))))
            fp.write("\n")
    return path


def write_state_file(path, mu=(0.0, 0.5)):
    """`1 2 2 / 1 2 / 1 2 / mu0 mu1` (grammar: pair_table_ucgld.cpp:565-652)."""
    with open(path, "w") as fp:
        fp.write("1 2 2\n1 2\n1 2\n%r %r\n" % (float(mu[0]), float(mu[1])))
    return path


# single-type, 2-state maps in the 1-based array form the C-ABI takes
def single_type_maps(mu=(0.0, 0.5), mass=1.0):
    n_states = np.array([0, 2], np.int32)
    formal_from = np.array([[0, 0], [1, 2]], np.int32)
    chem_pot = np.array([0.0, mu[0], mu[1]])
    masses = np.array([0.0, mass, mass])
    return dict(n_actual=1, n_formal=2, n_states=n_states, formal_from_actual=formal_from,
                chem_pot=chem_pot, mass=masses)


def fixture_dir(base):
    os.makedirs(base, exist_ok=True)
    return base
