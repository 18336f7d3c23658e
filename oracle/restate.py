"""oracle/restate.py -- CPU restatement (numpy + plain Python loops) of the reference's hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg -- never by
lammps_le_b200/.  Every function follows the reference file:line it cites (paths under /root/reference).
Parity is PINNED: tests/test_oracle.py checks this file against
  * the reference's own unit-test vectors (unittest/force-styles/tests/mol-pair-lj_cut.yaml, bond-fene.yaml,
    bond-harmonic.yaml; extracted into tests/golden/ref_yaml_*.npz by oracle/extract_ref_yaml.py), and
  * outputs of the compiled reference itself (oracle/_ref, built by oracle/build_ref.py) stored in
    tests/golden/forces_chain.npz and tests/golden/le_trace_small.npz (USER-LE has no tests of its own in the
    reference: SURVEY.md section 8c).
All arrays are in tag order (row t-1 = atom with tag t), exactly what the reference's Atom class holds in the
1-rank `atom_modify sort 0 0` configuration of the parity protocol.
"""
import math

import numpy as np

BIG = 1.0e20
TWO_1_3 = 1.2599210498948732   # src/MOLECULE/bond_fene.cpp:22
NEIGHMASK = 0x3FFFFFFF
SBBITS = 30


# ------------------------------------------------------------------------------------------------------------
# RanMars -- src/random_mars.cpp:29-95
# ------------------------------------------------------------------------------------------------------------
class RanMars:
    def __init__(self, seed):
        if seed <= 0 or seed > 900000000:
            raise ValueError("Invalid seed for Marsaglia random # generator")
        u = [0.0] * 98
        ij = (seed - 1) // 30082
        kl = (seed - 1) - 30082 * ij
        i = (ij // 177) % 177 + 2
        j = ij % 177 + 2
        k = (kl // 169) % 178 + 1
        l = kl % 169
        for ii in range(1, 98):
            s, t = 0.0, 0.5
            for _ in range(24):
                m = ((i * j) % 179) * k % 179
                i, j, k = j, k, m
                l = (53 * l + 1) % 169
                if (l * m) % 64 >= 32:
                    s += t
                t *= 0.5
            u[ii] = s
        self.u = u
        self.c = 362436.0 / 16777216.0
        self.cd = 7654321.0 / 16777216.0
        self.cm = 16777213.0 / 16777216.0
        self.i97, self.j97 = 97, 33
        self.ncalls = -1          # the constructor's own uniform() is not counted
        self.uniform()

    def uniform(self):
        u = self.u
        uni = u[self.i97] - u[self.j97]
        if uni < 0.0:
            uni += 1.0
        u[self.i97] = uni
        self.i97 -= 1
        if self.i97 == 0:
            self.i97 = 97
        self.j97 -= 1
        if self.j97 == 0:
            self.j97 = 97
        self.c -= self.cd
        if self.c < 0.0:
            self.c += self.cm
        uni -= self.c
        if uni < 0.0:
            uni += 1.0
        self.ncalls += 1
        return uni

    def skip(self, n):
        for _ in range(n):
            self.uniform()
        return self

    def c24(self):
        """RanMars::c * 2^24, the counter the harness records (oracle/ref_harness.cpp rng_c)"""
        return int(round(self.c * 16777216.0))


# ------------------------------------------------------------------------------------------------------------
# geometry helpers
# ------------------------------------------------------------------------------------------------------------
def image_shift(xi, xj, L):
    """integer box shift s such that xj + s*L is the image of j closest to i (Domain::closest_image,
    src/domain.cpp; the ghost AtomVec::pack_border made, x + pbc*prd)"""
    return -np.rint((xj - xi) / L)


def pair_rsq(xi, xj, L):
    """squared distance between i and the closest image of j with the reference's operation order
    (delx = xtmp - x[j][0]; rsq = delx*delx + dely*dely + delz*delz; src/npair_half_bin_newton.cpp:98-102)"""
    s = image_shift(xi, xj, L)
    xjj = np.where(s != 0, xj + s * L, xj)
    d = xi - xjj
    return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2], d, s


# ------------------------------------------------------------------------------------------------------------
# angle cosine -- AngleCosine::compute src/MOLECULE/angle_cosine.cpp:47-140 (E = K (1 + cos theta))
# ------------------------------------------------------------------------------------------------------------
def angle_cosine(x, L, a1, a2, a3, atype, k_of_type):
    """forces / energy / virial of the angles (a1, a2, a3): 0-based indices, a2 the centre; atype 1-based; k_of_type {type: K}.
    Arms are taken to the closest image of the end atoms as seen from the centre (the reference's ghost atoms)."""
    n = len(x)
    f = np.zeros((n, 3))
    e = 0.0
    vir = np.zeros(6)
    for m in range(len(a1)):
        i1, i2, i3 = int(a1[m]), int(a2[m]), int(a3[m])
        _, d1, _ = pair_rsq(x[i1], x[i2], L)          # x[i1] - x[i2] with i2 shifted to i1's image: the arm del1
        _, d2, _ = pair_rsq(x[i3], x[i2], L)
        rsq1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]
        rsq2 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]
        r1, r2 = math.sqrt(rsq1), math.sqrt(rsq2)
        c = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2]
        c /= r1 * r2
        c = min(1.0, max(-1.0, c))
        kk = k_of_type[int(atype[m])]
        e += kk * (1.0 + c)
        a11 = kk * c / rsq1
        a12 = -kk / (r1 * r2)
        a22 = kk * c / rsq2
        f1 = a11 * d1 + a12 * d2
        f3 = a22 * d2 + a12 * d1
        f[i1] += f1
        f[i2] -= f1 + f3
        f[i3] += f3
        # ev_tally (src/angle.cpp:236-270): v = del1 * f1 + del2 * f3
        vir += np.array([d1[0] * f1[0] + d2[0] * f3[0], d1[1] * f1[1] + d2[1] * f3[1], d1[2] * f1[2] + d2[2] * f3[2],
                         d1[0] * f1[1] + d2[0] * f3[1], d1[0] * f1[2] + d2[0] * f3[2], d1[1] * f1[2] + d2[1] * f3[2]])
    return f, e, vir


# ------------------------------------------------------------------------------------------------------------
# pair lj/cut -- PairLJCut::init_one src/pair_lj_cut.cpp:512-535, compute :68-140
# ------------------------------------------------------------------------------------------------------------
def lj_coeffs(eps, sigma, rc, shift):
    lj1 = 48.0 * eps * sigma ** 12.0
    lj2 = 24.0 * eps * sigma ** 6.0
    lj3 = 4.0 * eps * sigma ** 12.0
    lj4 = 4.0 * eps * sigma ** 6.0
    off = 0.0
    if shift and rc > 0.0:
        ratio = sigma / rc
        off = 4.0 * eps * (ratio ** 12.0 - ratio ** 6.0)
    return dict(lj1=lj1, lj2=lj2, lj3=lj3, lj4=lj4, offset=off, cutsq=rc * rc)


def pair_lj_cut(x, L, pairs_i, pairs_j, which, co, special_lj=(1.0, 0.0, 1.0, 1.0), shift=None):
    """forces / energy / virial of a half pair list (0-based atom indices, `which` = special bits).
    co: dict of scalars or of per-pair arrays.  `shift` [P,3]: explicit periodic image of j per pair (the ghost the
    list entry points at); default = the closest image.  Returns f[N,3], evdwl, virial[6]."""
    n = len(x)
    f = np.zeros((n, 3))
    if len(pairs_i) == 0:
        return f, 0.0, np.zeros(6)
    if shift is None:
        rsq, d, _ = pair_rsq(x[pairs_i], x[pairs_j], L)
    else:
        d = x[pairs_i] - (x[pairs_j] + shift * L)
        rsq = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    factor = np.asarray(special_lj)[which]
    m = rsq < co["cutsq"]
    r2inv = 1.0 / np.where(m, rsq, 1.0)
    r6inv = r2inv * r2inv * r2inv
    forcelj = r6inv * (co["lj1"] * r6inv - co["lj2"])
    fpair = np.where(m, factor * forcelj * r2inv, 0.0)
    fv = d * fpair[:, None]
    np.add.at(f, pairs_i, fv)
    np.add.at(f, pairs_j, -fv)
    evdwl = np.where(m, factor * (r6inv * (co["lj3"] * r6inv - co["lj4"]) - co["offset"]), 0.0).sum()
    vir = np.array([(d[:, 0] * d[:, 0] * fpair).sum(), (d[:, 1] * d[:, 1] * fpair).sum(), (d[:, 2] * d[:, 2] * fpair).sum(),
                    (d[:, 0] * d[:, 1] * fpair).sum(), (d[:, 0] * d[:, 2] * fpair).sum(), (d[:, 1] * d[:, 2] * fpair).sum()])
    return f, float(evdwl), vir


# ------------------------------------------------------------------------------------------------------------
# bonds -- BondFENE::compute src/MOLECULE/bond_fene.cpp:52-128, BondHarmonic::compute bond_harmonic.cpp:48-100
# ------------------------------------------------------------------------------------------------------------
def bond_forces(x, L, b1, b2, btype, coeffs):
    """forces / energy / virial / #FENE warnings of unique bonds (0-based indices, btype 1-based).
    coeffs: {btype: ("fene", (k, r0, eps, sigma)) | ("harmonic", (k, r0))}.  With newton_bond off the reference
    visits a bond that straddles the periodic boundary twice and tallies half each time: the sum is this."""
    n = len(x)
    f = np.zeros((n, 3))
    e = 0.0
    vir = np.zeros(6)
    warn = 0
    for k in range(len(b1)):
        i, j = int(b1[k]), int(b2[k])
        style, p = coeffs[int(btype[k])]
        rsq, d, _ = pair_rsq(x[i], x[j], L)
        rsq = float(rsq)
        if style == "fene":
            kk, r0, eps, sigma = p
            r0sq = r0 * r0
            rlogarg = 1.0 - rsq / r0sq
            if rlogarg < 0.1:
                warn += 1
                if rlogarg <= -3.0:
                    raise RuntimeError("Bad FENE bond")
                rlogarg = 0.1
            fbond = -kk / rlogarg
            sr6 = 0.0
            if rsq < TWO_1_3 * sigma * sigma:
                sr2 = sigma * sigma / rsq
                sr6 = sr2 * sr2 * sr2
                fbond += 48.0 * eps * sr6 * (sr6 - 0.5) / rsq
            eb = -0.5 * kk * r0sq * math.log(rlogarg)
            if rsq < TWO_1_3 * sigma * sigma:
                eb += 4.0 * eps * sr6 * (sr6 - 1.0) + eps
        elif style == "harmonic":
            kk, r0 = p[0], p[1]
            r = math.sqrt(rsq)
            dr = r - r0
            rk = kk * dr
            fbond = -2.0 * rk / r if r > 0.0 else 0.0
            eb = rk * dr
        else:
            raise ValueError(style)
        f[i] += d * fbond
        f[j] -= d * fbond
        e += eb
        vir += np.array([d[0] * d[0], d[1] * d[1], d[2] * d[2], d[0] * d[1], d[0] * d[2], d[1] * d[2]]) * fbond
    return f, e, vir, warn


def unique_bonds(num_bond, bond_type, bond_atom):
    """each bond once (newton_bond off stores it on both atoms): rows taken from the lower tag"""
    b1, b2, bt = [], [], []
    for i in range(len(num_bond)):
        for m in range(num_bond[i]):
            p = int(bond_atom[i, m])
            if i + 1 < p:
                b1.append(i); b2.append(p - 1); bt.append(int(bond_type[i, m]))
    return np.array(b1, int), np.array(b2, int), np.array(bt, int)


# ------------------------------------------------------------------------------------------------------------
# special lists -- Special::build src/special.cpp:55-154 (+ dedup/combine :611-762)
# ------------------------------------------------------------------------------------------------------------
def special_build(num_bond, bond_atom, special_lj=(0.0, 0.0, 0.0)):
    """per atom three tiers as sets (1-2, 1-3, 1-4), each without self and without lower tiers.  Tiers whose weight
    (and every later one) is 1.0 are not built at all (special.cpp:92-101,111-121)"""
    n = len(num_bond)
    ntier = 1 if (special_lj[1] == 1.0 and special_lj[2] == 1.0) else 2 if special_lj[2] == 1.0 else 3
    one = [set(int(v) for v in bond_atom[i, :num_bond[i]]) for i in range(n)]
    out = []
    for i in range(n):
        s1 = set(one[i]) - {i + 1}
        s2 = set()
        for a in s1:
            s2 |= one[a - 1]
        s2 -= s1 | {i + 1}
        if ntier < 2:
            s2 = set()
        s3 = set()
        for a in s2:
            s3 |= one[a - 1]
        s3 -= s1 | s2 | {i + 1}
        if ntier < 3:
            s3 = set()
        out.append((frozenset(s1), frozenset(s2), frozenset(s3)))
    return out


def special_tiers(nspecial, special):
    out = []
    for t in range(len(nspecial)):
        n1, n2, n3 = (int(v) for v in nspecial[t])
        row = special[t]
        out.append((frozenset(row[:n1].tolist()), frozenset(row[n1:n2].tolist()), frozenset(row[n2:n3].tolist())))
    return out


# ------------------------------------------------------------------------------------------------------------
# neighbor list -- NBinStandard::setup_bins src/nbin_standard.cpp:53-186, NBin::coord2bin src/nbin.cpp:120-150,
# NPairHalfBinNewton::build src/npair_half_bin_newton.cpp:35-160, NStencilHalfBin3dNewton::create
# src/nstencil_half_bin_3d_newton.cpp:26-38, NPair::find_special src/npair.h:112-136
# ------------------------------------------------------------------------------------------------------------
def ref_bins(boxlo, boxhi, cutneighmax):
    L = boxhi - boxlo
    binsizeinv = 1.0 / (0.5 * cutneighmax)
    nb = np.maximum((L * binsizeinv).astype(int), 1)
    return nb, 1.0 / (L / nb)


def coord2bin(x, lo, hi, nb, inv):
    """per dimension; the mbinlo offset cancels in the bin differences used below"""
    out = np.empty(x.shape, dtype=np.int64)
    for d in range(3):
        v = x[..., d]
        a = ((v - hi[d]) * inv[d]).astype(np.int64) + nb[d]
        b = np.minimum(((v - lo[d]) * inv[d]).astype(np.int64), nb[d] - 1)
        c = ((v - lo[d]) * inv[d]).astype(np.int64) - 1
        out[..., d] = np.where(v >= hi[d], a, np.where(v >= lo[d], b, c))
    return out


def half_neighbor_list(x, boxlo, boxhi, cutneigh, nspecial, special, special_lj=(0.0, 1.0, 1.0)):
    """per atom (tag order) the SET of (partner tag | special bits) the reference's half list stores on it.
    Brute force over all pairs; cutoff test with the reference's arithmetic (rsq <= cutneighsq)."""
    n = len(x)
    L = boxhi - boxlo
    cutsq = cutneigh * cutneigh
    nb, inv = ref_bins(boxlo, boxhi, cutneigh)
    flags = [0 if v == 0.0 else 1 if v == 1.0 else 2 for v in special_lj]   # Neighbor::init src/neighbor.cpp:349-369
    rows = [set() for _ in range(n)]
    bins_own = coord2bin(x, boxlo, boxhi, nb, inv)
    for i in range(n - 1):
        xj = x[i + 1:]
        rsq, d, s = pair_rsq(x[i][None, :], xj, L)
        for jj in np.nonzero(rsq <= cutsq)[0]:
            j = i + 1 + int(jj)
            # special bits (find_special): 1-2 excluded when its weight is 0, kept plain when 1, bits otherwise
            which = 0
            n1, n2, n3 = (int(v) for v in nspecial[i])
            row = special[i, :n3]
            hit = np.nonzero(row == j + 1)[0]
            if len(hit):
                k = int(hit[0])
                tier = 0 if k < n1 else 1 if k < n2 else 2
                if flags[tier] == 0:
                    continue
                if flags[tier] == 2:
                    which = tier + 1
            ghost = bool((s[jj] != 0).any())
            # ownership seen from i (j = closest image of atom j+1) and, if that fails, from j
            xjj = x[j] + s[jj] * L
            bj = coord2bin(xjj[None, :], boxlo, boxhi, nb, inv)[0]
            db = bj - bins_own[i]
            if (db == 0).all():
                if not ghost:
                    on_i = True                                     # owned j later in the bin list (j > i)
                else:
                    xi = x[i]
                    on_i = not (xjj[2] < xi[2] or (xjj[2] == xi[2] and (xjj[1] < xi[1] or (xjj[1] == xi[1] and xjj[0] < xi[0]))))
            else:
                on_i = db[2] > 0 or (db[2] == 0 and (db[1] > 0 or (db[1] == 0 and db[0] > 0)))
            if on_i:
                rows[i].add((j + 1) | (which << SBBITS))
            else:
                rows[j].add((i + 1) | (which << SBBITS))
    return rows


# ------------------------------------------------------------------------------------------------------------
# bond list -- NTopoBondAll::build src/ntopo_bond_all.cpp:39-86 (newton_bond off)
# ------------------------------------------------------------------------------------------------------------
def bond_list(xhold, L, num_bond, bond_type, bond_atom):
    """rows (tag_i, tag_partner, type) in the reference's order; a bond whose partner's closest image is a periodic
    ghost is listed from both atoms (i < ghost index always)"""
    rows = []
    for i in range(len(num_bond)):
        for m in range(num_bond[i]):
            p = int(bond_atom[i, m])
            s = image_shift(xhold[i], xhold[p - 1], L)
            if (s != 0).any() or i + 1 < p:
                rows.append((i + 1, p, int(bond_type[i, m])))
    return np.array(rows, dtype=np.int32).reshape(-1, 3)


# ------------------------------------------------------------------------------------------------------------
# USER-LE: shared edits
# ------------------------------------------------------------------------------------------------------------
def _delete_bond_slot(S, i, ptag):
    """fix_extrusion.cpp:656-668 / fix_ex_unload.cpp:293-302"""
    nb = int(S["num_bond"][i])
    for m in range(nb):
        if S["bond_atom"][i, m] == ptag:
            for k in range(m, nb - 1):
                S["bond_atom"][i, k] = S["bond_atom"][i, k + 1]
                S["bond_type"][i, k] = S["bond_type"][i, k + 1]
            S["num_bond"][i] = nb - 1
            break


def _special_remove12(S, i, ptag):
    """fix_extrusion.cpp:673-683"""
    sl = S["special"][i]
    n1, n3 = int(S["nspecial"][i, 0]), int(S["nspecial"][i, 2])
    m = 0
    while m < n1 and sl[m] != ptag:
        m += 1
    while m < n3 - 1:
        sl[m] = sl[m + 1]
        m += 1
    S["nspecial"][i] -= 1


def _special_insert12(S, i, ptag):
    """fix_extrusion.cpp:748-771 / fix_ex_load.cpp:570-588"""
    sl = S["special"][i]
    n1, n2, n3 = (int(v) for v in S["nspecial"][i])
    m = n1
    while m < n3 and sl[m] != ptag:
        m += 1
    if m < n3:
        for k in range(m, n3 - 1):
            sl[k] = sl[k + 1]
        n3 -= 1
        if m < n2:
            n2 -= 1
    if n3 == S["special"].shape[1]:
        raise RuntimeError("New bond exceeded special list size")
    for k in range(n3, n1, -1):
        sl[k] = sl[k - 1]
    sl[n1] = ptag
    S["nspecial"][i] = (n1 + 1, n2 + 1, n3 + 1)


def _dedup(nstart, nstop, copy):
    """FixExtrusion::dedup fix_extrusion.cpp:1116-1135"""
    m = nstart
    while m < nstop:
        dup = False
        for i in range(m):
            if copy[i] == copy[m]:
                copy[m] = copy[nstop - 1]
                nstop -= 1
                dup = True
                break
        if not dup:
            m += 1
    return nstop


def _rebuild_special_one(S, m):
    """fix_extrusion.cpp:1045-1108"""
    sp, ns = S["special"], S["nspecial"]
    tagm = m + 1
    copy = [int(v) for v in sp[m, :ns[m, 0]]]
    cn1 = len(copy)
    for i in range(cn1):
        n = copy[i] - 1
        copy += [int(v) for v in sp[n, :ns[n, 0]] if v != tagm]
    cn2 = _dedup(cn1, len(copy), copy)
    del copy[cn2:]
    for i in range(cn1, cn2):
        n = copy[i] - 1
        copy += [int(v) for v in sp[n, :ns[n, 0]] if v != tagm]
    cn3 = _dedup(cn2, len(copy), copy)
    del copy[cn3:]
    if cn3 > sp.shape[1]:
        raise RuntimeError("Special list size exceeded in fix bond/create")
    ns[m] = (cn1, cn2, cn3)
    sp[m, :cn3] = copy


def _update_topology(S, broken, created):
    """fix_extrusion.cpp:924-1002 (identical in fix_ex_load.cpp:700-751 / fix_ex_unload.cpp:417-463)"""
    n = len(S["num_bond"])
    sp, ns = S["special"], S["nspecial"]
    for i in range(n):
        infl = False
        row = set(int(v) for v in sp[i, :ns[i, 2]])
        for a, b in broken:
            if i + 1 in (a, b) or (a in row and b in row):
                infl = True
                break
        if infl:
            _rebuild_special_one(S, i)
    for i in range(n):
        infl = False
        row = set(int(v) for v in sp[i, :ns[i, 1]])
        for a, b in created:
            if i + 1 in (a, b) or a in row or b in row:
                infl = True
                break
        if infl:
            _rebuild_special_one(S, i)


def _bondcount(S, btype, strict):
    bc = np.zeros(len(S["num_bond"]), int)
    for i in range(len(bc)):
        for m in range(S["num_bond"][i]):
            if S["bond_type"][i, m] == btype:
                bc[i] += 1
                if strict and bc[i] > 1:
                    raise RuntimeError("Fix extrusion, more than one bond type 2")
    return bc


def copy_state(rec):
    S = {k: np.array(rec[k]).copy() for k in ("x", "xhold", "type", "num_bond", "bond_type", "bond_atom", "nspecial", "special")}
    S["L"] = np.asarray(rec["boxhi"]) - np.asarray(rec["boxlo"])
    S["boxlo"], S["boxhi"] = np.asarray(rec["boxlo"]), np.asarray(rec["boxhi"])
    return S


# ------------------------------------------------------------------------------------------------------------
# fix extrusion -- FixExtrusion::post_integrate src/USER-LE/fix_extrusion.cpp:256-872
# ------------------------------------------------------------------------------------------------------------
def fix_extrusion(S, rng, neutral, left, right, p_through, btype, roadblock, bondlist=None):
    """one slide step on state S (modified in place); returns (breakcount, createcount)"""
    n = len(S["num_bond"])
    x, typ, nb = S["x"], S["type"], S["num_bond"]
    bc = _bondcount(S, btype, True)
    to_remove = np.zeros(n + 2, int); to_add = np.zeros(n + 2, int)
    distsq = np.full(n + 2, BIG)
    if bondlist is None:
        bondlist = bond_list(S["xhold"], S["L"], nb, S["bond_type"], S["bond_atom"])

    def idx(tag):          # atom->map(tag) in the oracle configuration; out-of-range tags behave as "no such bead"
        return tag - 1 if 1 <= tag <= n else None

    def rsq_of(a, b):      # owned copies, raw coordinates (not minimum image): :431-434
        d = x[a] - x[b]
        return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]

    def side_ok(bead, barrier):
        """the && chain of :406-416 / :420-429 with its short-circuit draw consumption"""
        k = idx(bead)
        if k is None:
            return False
        if not (nb[k] - bc[k] == 2 and bc[k] == 0):
            return False
        ty = typ[k]
        if ty not in (left, right, roadblock, neutral):
            return False
        if ty == barrier and not (p_through > rng.uniform()):
            return False
        if ty == roadblock and not (p_through > rng.uniform()):
            return False
        return True

    for t1, t2, ty in bondlist:
        if ty != btype:
            continue
        a, b = (int(t1), int(t2)) if t1 < t2 else (int(t2), int(t1))
        i1, i2 = a - 1, b - 1
        if nb[i1] in (0, 1) or nb[i2] in (0, 1) or bc[i1] != 1 or bc[i2] != 1:
            continue
        if side_ok(a - 1, left):
            Lk = a - 2
            if side_ok(b + 1, right):
                Rk = b
                rsq = rsq_of(Lk, Rk)
                if rsq >= distsq[Lk] and rsq >= distsq[Rk]:
                    continue
                if rsq < distsq[Lk]:
                    distsq[Lk] = rsq; to_add[Lk] = b + 1
                if rsq < distsq[Rk]:
                    distsq[Rk] = rsq; to_add[Rk] = a - 1
            else:
                rsq = rsq_of(Lk, i2)
                if rsq >= distsq[Lk]:
                    continue
                distsq[Lk] = rsq; to_add[Lk] = b
                if distsq[i2] == BIG:
                    distsq[i2] = rsq; to_add[i2] = a - 1
            to_remove[i1] = b; to_remove[i2] = a
        elif side_ok(b + 1, right):
            Rk = b
            rsq = rsq_of(i1, Rk)
            if rsq >= distsq[Rk]:
                continue
            if distsq[i1] == BIG:
                distsq[i1] = rsq; to_add[i1] = b + 1
            distsq[Rk] = rsq; to_add[Rk] = a
            to_remove[i1] = b; to_remove[i2] = a

    # reconciliation :517-599
    for i in range(n):
        if to_add[i] == 0:
            continue
        j = to_add[i] - 1
        if to_add[j] == i + 1:
            continue
        ti, tj = i + 1, j + 1
        if ti < tj:
            lb, rb = ti, tj - 2          # 0-based indices of tags ti+1, tj-1
            opts = [(lb, rb), (i, rb), (lb, j), (i, j)]
        else:
            lb, rb = tj, ti - 2          # tags tj+1, ti-1
            opts = [(lb, rb), (i, lb), (rb, j), (i, j)]
        for p, q in opts:
            if p + 1 == to_remove[q] and to_remove[p] == q + 1:
                to_remove[p] = 0; to_remove[q] = 0
                break

    def ta(tag):
        return to_add[tag - 1] if 1 <= tag <= n else 0

    def tr(tag):
        return to_remove[tag - 1] if 1 <= tag <= n else 0

    final_remove = np.zeros(n, int); final_add = np.zeros(n, int)
    nbreak = 0
    for i in range(n):                    # break loop :618-692
        if to_remove[i] == 0:
            continue
        j = to_remove[i] - 1
        if to_remove[j] != i + 1:
            continue
        lb, rb = min(i + 1, j + 1), max(i + 1, j + 1)
        if ta(lb - 1) == rb and ta(rb) == lb - 1 and ta(lb) == rb + 1 and ta(rb + 1) == lb:
            to_add[lb - 2] = rb + 1; to_add[rb] = lb - 1; to_add[lb - 1] = 0; to_add[rb - 1] = 0
        if (ta(lb - 1) == rb and ta(rb) == lb - 1) or (ta(lb - 1) == rb + 1 and ta(rb + 1) == lb - 1) or \
                (ta(lb) == rb + 1 and ta(rb + 1) == lb):
            _delete_bond_slot(S, i, j + 1)
            _special_remove12(S, i, j + 1)
            final_remove[i] = j + 1; final_remove[j] = i + 1
            if i < j:
                nbreak += 1
    ncreate = 0
    bpa = S["bond_type"].shape[1]
    for i in range(n):                    # create loop :699-786
        if to_add[i] == 0:
            continue
        j = to_add[i] - 1
        if to_add[j] != i + 1:
            continue
        if nb[i] == bpa:
            continue
        lb, rb = min(i + 1, j + 1), max(i + 1, j + 1)
        if (tr(lb + 1) == rb and tr(rb) == lb + 1) or (tr(lb + 1) == rb - 1 and tr(rb - 1) == lb + 1) or \
                (tr(lb) == rb - 1 and tr(rb - 1) == lb):
            S["bond_type"][i, nb[i]] = btype
            S["bond_atom"][i, nb[i]] = j + 1
            nb[i] += 1
            _special_insert12(S, i, j + 1)
            final_add[i] = j + 1; final_add[j] = i + 1
            if i < j:
                ncreate += 1
    if nbreak == 0 and ncreate == 0:
        return 0, 0
    if nbreak != ncreate:
        raise RuntimeError("Numbers of created and broken bonds are not equal")
    broken = [(i + 1, int(final_remove[i])) for i in range(n) if final_remove[i] and i + 1 < final_remove[i]]
    created = [(i + 1, int(final_add[i])) for i in range(n) if final_add[i] and i + 1 < final_add[i]]
    _update_topology(S, broken, created)
    return nbreak, ncreate


# ------------------------------------------------------------------------------------------------------------
# fix ex_unload -- FixExUnload::post_integrate src/USER-LE/fix_ex_unload.cpp:172-372
# ------------------------------------------------------------------------------------------------------------
def fix_ex_unload(S, rng, btype, rc, fraction, bondlist=None):
    n = len(S["num_bond"])
    x, L = S["x"], S["L"]
    cutsq = rc * rc
    partner = np.zeros(n, int); distsq = np.zeros(n)
    if bondlist is None:
        bondlist = bond_list(S["xhold"], L, S["num_bond"], S["bond_type"], S["bond_atom"])
    for t1, t2, ty in bondlist:
        if ty != btype:
            continue
        i1, i2 = int(t1) - 1, int(t2) - 1
        # i2 is a periodic ghost iff its closest image at the last rebuild was shifted; the ghost follows its owner
        # (Comm::forward_comm, x + pbc*prd) and has its own scratch slot, overwritten by forward_comm_fix
        s = image_shift(S["xhold"][i1], S["xhold"][i2], L)
        ghost = bool((s != 0).any())
        xj = x[i2] + s * L if ghost else x[i2]
        d = x[i1] - xj
        rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        if rsq <= cutsq:
            continue
        if rsq > distsq[i1]:
            partner[i1] = t2; distsq[i1] = rsq
        if not ghost and rsq > distsq[i2]:
            partner[i2] = t1; distsq[i2] = rsq
    prob = np.zeros(n)
    if fraction < 1.0:
        for i in range(n):
            if partner[i]:
                prob[i] = rng.uniform()
    final = np.zeros(n, int)
    nbreak = 0
    for i in range(n):
        if partner[i] == 0:
            continue
        j = partner[i] - 1
        if partner[j] != i + 1:
            continue
        if fraction < 1.0:
            if (prob[i] if i < j else prob[j]) >= fraction:
                continue
        _delete_bond_slot(S, i, j + 1)
        _special_remove12(S, i, j + 1)
        final[i] = j + 1; final[j] = i + 1
        if i < j:
            nbreak += 1
    if nbreak:
        broken = [(i + 1, int(final[i])) for i in range(n) if final[i] and i + 1 < final[i]]
        _update_topology(S, broken, [])
    return nbreak


# ------------------------------------------------------------------------------------------------------------
# fix ex_load -- FixExLoad::post_integrate src/USER-LE/fix_ex_load.cpp:329-655
# ------------------------------------------------------------------------------------------------------------
def fix_ex_load(S, rng, itype, jtype, rc, btype, fraction, imax, inew, jmax, jnew, neigh_offsets, neigh_entries, ancestor=False, bc=None):
    """neigh_*: the half pair list of the last rebuild in the reference's order (CSR, tag order, entries =
    partner tag | special bits).
    ancestor=True: fix bond/create (src/MC/fix_bond_create.cpp:349-629), the fix ex_load was derived from -- the same loops
    without the loop-extrusion rules of fix_ex_load.cpp:470-484 (so a periodic ghost can be a partner: its coordinate is the
    owned atom's plus the shift of the image found at the last rebuild), with `bc`, the bond counts the fix took in the setup
    of its first run (fix_bond_create.cpp:302-345) and has only added its own creations to since (:576), updated in place"""
    n = len(S["num_bond"])
    x, L, typ, nb = S["x"], S["L"], S["type"], S["num_bond"]
    cutsq = rc * rc
    if bc is None:
        bc = _bondcount(S, btype, False)
    partner = np.zeros(n + 2, int); distsq = np.full(n + 2, BIG)
    for i in range(n):
        it = typ[i]
        for e in neigh_entries[neigh_offsets[i]:neigh_offsets[i + 1]]:
            j = (int(e) & NEIGHMASK) - 1
            s = image_shift(S["xhold"][i], S["xhold"][j], L)
            ghost = bool((s != 0).any())
            jt = typ[j]
            possible = False
            if it == itype and jt == jtype:
                possible = (imax == 0 or bc[i] < imax) and (jmax == 0 or bc[j] < jmax)
            elif it == jtype and jt == itype:
                possible = (jmax == 0 or bc[i] < jmax) and (imax == 0 or bc[j] < imax)
            if not possible:
                continue
            if not ancestor:
                if abs(i - j) != 2:
                    continue
                mid = (i + j) // 2
                if partner[mid] != 0:
                    continue
                if nb[i] != 2:
                    continue
                if ghost or nb[j] != 2:       # a ghost's num_bond is never communicated and reads 0 (atom_vec_bond.cpp:41)
                    continue
                if nb[mid] != 2:
                    continue
            if (S["special"][i, :S["nspecial"][i, 0]] == j + 1).any():
                continue
            d = x[i] - (x[j] + s * L if (ancestor and ghost) else x[j])
            rsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
            if rsq >= cutsq:
                continue
            if rsq < distsq[i]:
                partner[i] = j + 1; distsq[i] = rsq
            if rsq < distsq[j]:
                partner[j] = i + 1; distsq[j] = rsq
    prob = np.zeros(n)
    if fraction < 1.0:
        for i in range(n):
            if partner[i]:
                prob[i] = rng.uniform()
    final = np.zeros(n, int)
    ncreate = 0
    bpa = S["bond_type"].shape[1]
    for i in range(n):
        if partner[i] == 0:
            continue
        j = partner[i] - 1
        if partner[j] != i + 1:
            continue
        if fraction < 1.0:
            if (prob[i] if i < j else prob[j]) >= fraction:
                continue
        if nb[i] == bpa:
            raise RuntimeError("New bond exceeded bonds per atom in fix %s" % ("bond/create" if ancestor else "ex_load"))
        S["bond_type"][i, nb[i]] = btype
        S["bond_atom"][i, nb[i]] = j + 1
        nb[i] += 1
        _special_insert12(S, i, j + 1)
        bc[i] += 1
        if typ[i] == itype:
            if bc[i] == imax:
                typ[i] = inew
        else:
            if bc[i] == jmax:
                typ[i] = jnew
        final[i] = j + 1; final[j] = i + 1
        if i < j:
            ncreate += 1
    if ncreate:
        created = [(i + 1, int(final[i])) for i in range(n) if final[i] and i + 1 < final[i]]
        _update_topology(S, [], created)
    return ncreate
