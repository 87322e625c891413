"""fix cluster_switch on the GPU against the reference's own compiled source (oracle/_ref): cluster
labels, mol_state, accept decisions (same RanPark stream, bit for bit), atom types and MC
statistics.  Deck: pair_style table_rleucg_interface with 4 state types — 1 = ON (two-state density
type, substate type 2), 3 = OFF, 4 = the inert partner molecules — molecules of 4 sites; switchable
molecule m is tied to its partner m - mol_offset."""
import os

import numpy as np
import pytest

import ref_binding as rb
from decks import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built")]

RLE_STATE = "3 4\n2 density use_entropy\n12.0 1.5\n0.3\n1 density no_entropy\n1 density no_entropy\n"
CONTACTS = [(1, 1), (1, 3), (3, 1), (3, 3)]
PAIRS = [(1, 1, "UCG_00"), (1, 2, "UCG_01"), (1, 3, "UCG_01"), (1, 4, "UCG_01"), (2, 2, "UCG_11"), (2, 3, "UCG_01"),
         (2, 4, "UCG_01"), (3, 3, "UCG_01"), (3, 4, "UCG_01"), (4, 4, "UCG_01")]


def _system(ncell):
    from lammps_ucg_dev_b200 import synth
    liq = synth.fcc_liquid(ncell, mol_size=4)
    nmol = liq.n // 4
    half = nmol // 2
    liq.type[:] = 4
    sw = liq.molecule > half                                  # switchable molecules half+1 .. nmol
    liq.type[sw] = np.where((liq.molecule[sw] % 3) == 0, 3, 1)   # every third one starts OFF
    return liq, half


def _ref(liq, half, tmp_path, fixtures, cutoff, freq, seed, prob_on):
    sf = tmp_path / "rle.conf"
    sf.write_text(RLE_STATE)
    (tmp_path / "rates.txt").write_text(f"{prob_on}\n1\n1\n3\n")
    (tmp_path / "contacts.txt").write_text("ncontacts 2\natomspercontact 2\n" + "".join(f"{a} {b}\n" for a, b in CONTACTS))
    t = fixtures["table4096"]
    s = rb.RefSim()
    s.box(liq.box_lo, liq.box_hi, 4)
    s.atoms(liq)
    cmds = ["newton off", "neighbor 0.3 bin", "timestep 0.002", f"pair_style table_rleucg_interface linear 4096 {sf}"]
    cmds += [f"pair_coeff {i} {j} {t} {kw} 2.5" for i, j, kw in PAIRS]
    cmds += ["fix 0 all ttarget/stub 1.0", "fix 1 all nve/ucgld/wall/hard",
             f"fix 2 all cluster_switch {half + 1} {half} {cutoff} {seed} rateFreq {freq} rateFile {tmp_path}/rates.txt "
             f"contactFile {tmp_path}/contacts.txt"]
    cwd = os.getcwd()
    os.chdir(tmp_path)            # the fix writes cluster_assignment.log / state_assignment.log into the cwd
    try:
        for c in cmds:
            s.command(c)
    finally:
        os.chdir(cwd)
    return s


def _gpu(pkg, liq, half, fixtures, cutoff, freq, seed, prob_on):
    from lammps_ucg_dev_b200 import engine
    t = fixtures["table4096"]
    ctx = pkg.Context(0)
    ctx.set_units(1.0, 1.0, 1.0)
    ctx.set_box(liq.box_lo, liq.box_hi)
    ctx.set_timestep(0.002)
    idx = {kw: engine.HostTable.from_file(t, kw, 2.5, 1, 4096).upload(ctx) for kw in ("UCG_00", "UCG_01", "UCG_11")}
    tabindex = np.zeros((5, 5), np.int32)
    for i, j, kw in PAIRS:
        tabindex[i, j] = tabindex[j, i] = idx[kw]
    cutsq = np.zeros((5, 5)); cutsq[1:, 1:] = 2.5 ** 2
    ctx.pair_rleucg_configure(4, [0, 1, 1, 2, 3], 3, [0, 2, 1, 1], [0, 1, 0, 0], [0.0, 12.0, 0.0, 0.0], [0.0, 1.5, 0.0, 0.0],
                              [0.0, 0.3, 0.0, 0.0, 0.0], tabindex, cutsq, [0.0, 1.0, 1.0, 1.0, 1.0], 1.0)
    ctx.neigh_configure(0.3)
    engine.upload_liquid(ctx, liq)
    ctx.cluster_configure(half + 1, half, cutoff, seed, prob_on, [1], [3], CONTACTS, 4)
    ctx.deck_configure(pair_style=2, nve=2, thermo_every=0, cluster_freq=freq)
    return ctx


def _last_log_row(path):
    rows = [r for r in open(path).read().strip().split("\n") if r.strip()]
    return np.array(rows[-1].split()[1:], dtype=np.int64), len(rows)


@pytest.mark.parametrize("cutoff,prob_on", [(1.08, 0.3), (1.10, 0.7)])
def test_cluster_switch_trajectory(pkg, fixtures, tmp_path, cutoff, prob_on):
    liq, half = _system(6)
    freq, seed, nsteps = 5, 15123, 23          # checks at steps 1, 6, 11, 16, 21
    ref = _ref(liq, half, tmp_path, fixtures, cutoff, freq, seed, prob_on)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.setup(1)
        ref.run(nsteps, 1)        # eflag on every step: the shipped pair style feeds a stale evdwl otherwise (Q16)
    finally:
        os.chdir(cwd)
    a = ref.get_atoms()
    ctx = _gpu(pkg, liq, half, fixtures, cutoff, freq, seed, prob_on)
    ctx.setup()
    ctx.run(nsteps)
    b = ctx.atoms_download(["x", "f", "type"])
    # accept decisions -> atom types: exact
    assert np.array_equal(a["type"], b["type"])
    changed = (a["type"] != liq.type).sum()
    assert changed > 0, "the deck must actually switch something"
    # MC statistics and cluster size (compute_vector)
    st = ctx.cluster_stats()
    refst = np.array([ref.fix_vector(2, k) for k in range(7)])
    assert np.array_equal(st[:7], refst), (st, refst)
    assert 1 < refst[6] < 2 * half, "cluster neither trivial nor everything"
    # per-molecule cluster membership and state of the last check (the reference's log files)
    g = ctx.cluster_get()
    flags, nrows = _last_log_row(tmp_path / "cluster_assignment.log")
    states, _ = _last_log_row(tmp_path / "state_assignment.log")
    assert nrows == 5
    cid = g["mol_cluster"][half + 1]
    assert np.array_equal((g["mol_cluster"] == cid).astype(np.int64), flags)
    # state_assignment.log is written by check_cluster, i.e. before that step's switching: undo our last flips
    st_before = g["mol_state"].copy()
    acc = g["mol_accept"] == 1
    st_before[acc] = 1 - st_before[acc]
    assert np.array_equal(st_before.astype(np.int64), states)
    # and the dynamics with the switched types
    box = liq.box_hi - liq.box_lo
    dx = b["x"] - a["x"]
    dx -= box * np.round(dx / box)
    assert np.abs(dx).max() <= 1e-9
    assert rel_err(b["f"], a["f"]) <= 1e-6


def test_cluster_labels_direct(pkg, fixtures, tmp_path):
    """check_cluster alone against a union-find over brute-force contacts"""
    liq, half = _system(6)
    cutoff = 1.09
    ctx = _gpu(pkg, liq, half, fixtures, cutoff, 5, 7, 0.5)
    ctx.neigh_build()
    ncl = ctx.cluster_check()
    g = ctx.cluster_get()
    # brute force: contacts between switchable sites (types 1, 3) of different molecules, ties (m, m-half)
    box = liq.box_hi - liq.box_lo
    sw = np.where(liq.type != 4)[0]
    parent = np.arange(2 * half + 2)
    init = np.arange(2 * half + 2)
    for m in range(half + 1, 2 * half + 1):
        init[m - half] = m

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for m in range(half + 1, 2 * half + 1):
        parent[find(m - half)] = find(m)
    xs = liq.x[sw]
    for k, i in enumerate(sw):
        d = xs - liq.x[i]
        d -= box * np.round(d / box)
        near = np.where((d * d).sum(1) < cutoff * cutoff)[0]
        for q in near:
            j = sw[q]
            if liq.molecule[j] != liq.molecule[i]:
                parent[find(liq.molecule[i])] = find(liq.molecule[j])
    roots = np.array([find(m) for m in range(2 * half + 2)])
    expect = np.full(2 * half + 2, -1)
    for r in np.unique(roots[1:2 * half + 1]):
        members = np.where(roots == r)[0]
        members = members[(members >= 1) & (members <= 2 * half)]
        expect[members] = init[members].min()
    assert np.array_equal(g["mol_cluster"][1:2 * half + 1], expect[1:2 * half + 1])
    assert ncl == (expect[1:] == expect[half + 1]).sum()
