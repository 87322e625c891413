"""Host-side logic of the N>1 path on CPU: brick assignment, and the exchange orchestration
(BrickCluster + DistTransport) under torch.distributed with the gloo backend, world_size 2.
The bricks here are a tiny numpy model of the halo protocol (1-D ring of sites); the GPU
kernels themselves are covered by tests/test_gpu_multibrick.py."""
import ctypes
import os
import socket

import numpy as np
import pytest


def test_procgrid_and_brick_assignment(pkg):
    from lammps_ucg_dev_b200 import multigpu
    assert multigpu.procgrid_for(1) == (1, 1, 1)
    assert multigpu.procgrid_for(2) == (2, 1, 1)
    assert multigpu.procgrid_for(4) == (2, 2, 1)
    assert multigpu.procgrid_for(8) == (2, 2, 2)
    rng = np.random.default_rng(1)
    lo, hi = np.zeros(3), np.array([10.0, 7.0, 5.0])
    x = rng.uniform(lo, hi, (5000, 3))
    x[0] = (5.0, 3.5, 2.5)          # exactly on every split plane -> upper brick
    x[1] = (np.nextafter(5.0, 0), 0, 0)
    r = multigpu.brick_of(x, lo, hi, (2, 2, 2))
    assert r[0] == 7 and r[1] == 0
    assert set(np.unique(r)) == set(range(8))
    cx = (x[:, 0] >= 5.0).astype(int); cy = (x[:, 1] >= 3.5).astype(int); cz = (x[:, 2] >= 2.5).astype(int)
    assert np.array_equal(r, (cz * 2 + cy) * 2 + cx)


REC = dict(migrate=16, border=16, forward=16)
L, CUT = 16.0, 1.5


class RingBrick:
    """sites on a periodic ring [0,L); brick r owns [r*L/n, (r+1)*L/n); records = (tag, x) doubles"""

    def __init__(self, rank, nranks, tags, x):
        self.rank, self.n = rank, nranks
        self.lo, self.hi = rank * L / nranks, (rank + 1) * L / nranks
        self.tags, self.x = np.array(tags, float), np.array(x, float)
        self.ghosts = {}

    @staticmethod
    def _view(ptr, nrec):
        if nrec == 0:
            return np.zeros((0, 2))
        buf = (ctypes.c_double * (2 * nrec)).from_address(ptr)
        return np.ctypeslib.as_array(buf).reshape(nrec, 2)

    def migrate_prepare(self):
        self.x %= L
        self.dest = np.minimum((self.x / (L / self.n)).astype(int), self.n - 1)
        return np.bincount(self.dest[self.dest != self.rank], minlength=self.n).astype(np.int32)

    def migrate_pack(self, ptr):
        leave = self.dest != self.rank
        order = np.argsort(self.dest[leave], kind="stable")
        out = self._view(ptr, int(leave.sum()))
        out[:, 0] = self.tags[leave][order]; out[:, 1] = self.x[leave][order]
        self.tags, self.x = self.tags[~leave], self.x[~leave]

    def migrate_unpack(self, ptr, nrecv):
        rec = self._view(ptr, nrecv)
        self.tags = np.concatenate([self.tags, rec[:, 0]]); self.x = np.concatenate([self.x, rec[:, 1]])

    def neigh_build_local(self):
        left, right = (self.rank - 1) % self.n, (self.rank + 1) % self.n
        sends = []
        for i in np.nonzero(self.x <= self.lo + CUT)[0]:
            sends.append((left, i, L if self.rank == 0 else 0.0))
        for i in np.nonzero(self.x >= self.hi - CUT)[0]:
            sends.append((right, i, -L if self.rank == self.n - 1 else 0.0))
        sends.sort(key=lambda s: s[0])
        self.sends = sends

    def halo_send_counts(self):
        return np.bincount([s[0] for s in self.sends], minlength=self.n).astype(np.int32)

    def _pack(self, ptr):
        out = self._view(ptr, len(self.sends))
        for k, (_, i, sh) in enumerate(self.sends):
            out[k] = (self.tags[i], self.x[i] + sh)

    halo_pack_border = _pack
    halo_pack_forward = _pack

    def halo_unpack_border(self, ptr, recv_counts):
        self.nrecv = int(np.sum(recv_counts))
        self.ghosts = {int(t): xx for t, xx in self._view(ptr, self.nrecv)}

    def halo_unpack_forward(self, ptr):
        self.ghosts = {int(t): xx for t, xx in self._view(ptr, self.nrecv)}

    def neigh_build_finish(self):
        self.xhold = self.x.copy()

    def ghosts_forward(self):
        pass

    def neigh_decide_local(self):
        return int(np.any(np.abs(self.x - self.xhold) > 0.25))


def _expected_ghosts(all_tags, all_x, lo, hi):
    exp = {}
    for t, x in zip(all_tags, all_x):
        for sh in (-L, 0.0, L):
            xs = x + sh
            if not (lo <= x < hi and sh == 0.0) and (lo - CUT <= xs < lo or hi <= xs <= hi + CUT):
                exp[int(t)] = xs
    return exp


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.load_package()
    from lammps_ucg_dev_b200 import multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    n = 200
    tags, x = np.arange(1, n + 1), rng.uniform(0, L, n)
    mine = (np.arange(n) % world) == rank               # deliberately wrong owners: migration must fix it
    brick = RingBrick(rank, world, tags[mine], x[mine])
    alloc = lambda r, nbytes: torch.zeros(max(int(nbytes), 16), dtype=torch.uint8)
    cl = multigpu.BrickCluster({rank: brick}, multigpu.DistTransport(dist, torch.device("cpu")), alloc, REC)
    cl.rebuild()
    ok = True
    drift = rng.uniform(-0.02, 0.02, n)
    dmap = dict(zip(tags, drift))
    for step in range(40):
        x = (x + drift) % L                              # global truth (wrapped)
        # the bricks integrate unwrapped, like the MD step does between rebuilds
        brick.x = brick.x + np.array([dmap[int(t)] for t in brick.tags])
        if cl.decide():
            cl.rebuild()
        else:
            cl.forward()
        exp = _expected_ghosts(tags, x, brick.lo, brick.hi)
        known = dict(brick.ghosts)
        # sites keep their owner between rebuilds: a site that just crossed a face is still "mine"
        known.update({int(t): xx for t, xx in zip(brick.tags, brick.x)})
        # every site within CUT (minus the skin) of my faces must be known to me at the right periodic image
        need = {t: xs for t, xs in exp.items()
                if (brick.lo - CUT + 0.3 <= xs < brick.lo) or (brick.hi <= xs <= brick.hi + CUT - 0.3)}
        for t, xs in need.items():
            if t not in known or abs(known[t] - xs) > 1e-9:
                ok = False
    total = torch.tensor([len(brick.tags)])
    dist.all_reduce(total)
    q.put((rank, ok, int(total.item()), cl.nrebuilds))
    dist.destroy_process_group()


def test_exchange_orchestration_gloo_world2(pkg):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert all(r[2] == 200 for r in res)
    assert res[0][3] == res[1][3] and res[0][3] >= 2
