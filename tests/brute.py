"""Independent O(N^2) numpy restatement of the UCG-LD pair formulas (no neighbor list, minimum
image) used to cross-check the oracle on small systems (last row)."""
import numpy as np


def linear_table(tab, rsq):
    """tab: dict(innersq, delta, invdelta, e, f) -> (u, f/r) by the LINEAR rule
    (pair_table_ucgld.cpp:446-453, 476-477)."""
    it = ((rsq - tab["innersq"]) * tab["invdelta"]).astype(np.int64)
    rsq_it = tab["innersq"] + it * tab["delta"]
    frac = (rsq - rsq_it) * tab["invdelta"]
    e, f = tab["e"], tab["f"]
    return e[it] + frac * (e[it + 1] - e[it]), f[it] + frac * (f[it + 1] - f[it])


def ucgld_bruteforce(x, box, lam, state, tabs, cutsq, mu, kT):
    """tabs[(a,b)] for a,b in {0,1}.  Returns f (n,3), ucgforce, scores (n,2), E, virial(6)."""
    n = x.shape[0]
    f = np.zeros((n, 3))
    uf = np.full(n, -(mu[1] - mu[0]))
    sc = np.zeros((n, 2))
    sc[:, 1] = -(mu[1] - mu[0]) / kT
    E = 0.0
    vir = np.zeros(6)
    for i in range(n):
        d = x[i] - x
        d -= box * np.round(d / box)
        rsq = (d * d).sum(1)
        m = (rsq < cutsq) & (np.arange(n) != i)
        j = np.nonzero(m)[0]
        dj, r2 = d[j], rsq[j]
        u = {}
        fp = {}
        for a in (0, 1):
            for b in (0, 1):
                u[a, b], fp[a, b] = linear_table(tabs[a, b], r2)
        li, lj = lam[i], lam[j]
        e = (1 - li) * (1 - lj) * u[0, 0] + (1 - li) * lj * u[0, 1] + (1 - lj) * li * u[1, 0] + li * lj * u[1, 1]
        fpair = (1 - li) * (1 - lj) * fp[0, 0] + (1 - li) * lj * fp[0, 1] + (1 - lj) * li * fp[1, 0] + li * lj * fp[1, 1]
        f[i] = (dj * fpair[:, None]).sum(0)
        uf[i] -= (lj * (u[1, 1] - u[0, 1]) + (1 - lj) * (u[1, 0] - u[0, 0])).sum()
        sj = state[j]
        sc[i, 0] -= np.where(sj == 1, u[0, 1], u[0, 0]).sum() / kT
        sc[i, 1] -= np.where(sj == 1, u[1, 1], u[1, 0]).sum() / kT
        E += 0.5 * e.sum()
        w = 0.5 * fpair
        vir += np.array([(dj[:, 0] ** 2 * w).sum(), (dj[:, 1] ** 2 * w).sum(), (dj[:, 2] ** 2 * w).sum(),
                         (dj[:, 0] * dj[:, 1] * w).sum(), (dj[:, 0] * dj[:, 2] * w).sum(),
                         (dj[:, 1] * dj[:, 2] * w).sum()])
    return f, uf, sc, E, vir
