"""Explicit-solution builder on top of the batched implicit solver.

The reference hands the condensed mpQP to ParametricDAQP.jl (`mpsolve` + `build_tree`,
/root/reference/src/explicit.jl:25-52,65-71), which explores critical regions one QP at a time.
Here the GPU does the exploring: the implicit kernel solves a few million sampled parameters in
milliseconds and the distinct optimal active sets it returns ARE the critical regions that have
volume in the range; the set is then closed under facet adjacency (sets one row away from a known
region, kept when their Chebyshev LP says they are full-dimensional), which catches regions too
thin for any sample; each set gets its exact polyhedron and affine law (the usual mp-QP
algebra on the single-sided form of `make_singlesided`, /root/reference/src/utils.jl:329-391), a
Chebyshev LP drops the lower-dimensional ones, and a binary search tree over region facets is
built in the node-array layout `lmpc_gpu_create_bst` takes.  Laws and half-planes are exact;
`verify` reports how the tree compares with the implicit solve on fresh parameters.

Soft rows are made explicit the way the reference does it (`make_singlesided`, utils.jl:345-380): one
slack variable per soft row with weight soft_weight, entering the row with its R^-1-norm, so that AFTI-16
can be made explicit too.  With `clipping` (the reference's default, `build_tree!(...; clipping=true)`,
src/explicit.jl:65) a leaf only has to reproduce the control AFTER clipping to the input bounds: regions
where an output sits at its bound are absorbed by a neighbouring unsaturated law whenever that law,
extended, lies beyond the bound on them (an LP per region decides), which is what shrinks the invpend
tree from thousands of nodes to a few hundred.  `load_pdaqp_codegen` reads the arrays that
`ParametricDAQP.codegen` emits (src/codegen.jl:62-66), so trees computed by the real package can be loaded
without Julia.  Offline code, numpy + scipy's LP, nothing here is on the evaluate path.
"""
from __future__ import annotations

import numpy as np

BIG = 1e20


def make_singlesided(q, soft_weight=1e6):
    """A z <= b + W theta from the two-sided mpQP dict of MPC.condense() (utils.jl:329-391, single_soft =
    false).  Soft rows get one slack variable each: z grows by n_soft columns, H by soft_weight * I, and the
    slack enters both sides of its row with coefficient -||A0_j R^-1|| (the row's norm in the LDP metric,
    which is what makes rho = 1 / soft_weight the regularisation of the NORMALISED row in the dual).
    Returns (H, f, f_theta, A, b, W, src) where src[i] = (two-sided row, +1 upper / -1 lower)."""
    H = np.asarray(q["H"], float)
    n = H.shape[0]
    f = np.asarray(q["f"], float).ravel()
    fth = np.asarray(q["f_theta"], float).reshape(n, -1)
    bu, bl = np.asarray(q["bu"], float).ravel(), np.asarray(q["bl"], float).ravel()
    m = len(bu)
    Ag = np.asarray(q["A"], float).reshape(-1, n)
    ms = m - Ag.shape[0]
    A0 = np.vstack([np.hstack([np.eye(ms), np.zeros((ms, n - ms))]), Ag])
    W0 = np.asarray(q["W"], float).reshape(m, -1)
    soft = (np.asarray(q["senses"]).astype(int) & 8) != 0
    if np.any(soft):
        ids = np.where(soft)[0]
        R = np.linalg.cholesky((H + H.T) / 2).T                      # H = R'R, R upper
        nf = np.linalg.norm(np.linalg.solve(R.T, A0[ids].T), axis=0)  # ||A0_j R^-1||
        ns = len(ids)
        A0 = np.hstack([A0, np.zeros((m, ns))])
        A0[ids, n + np.arange(ns)] = -nf
        H = np.block([[H, np.zeros((n, ns))], [np.zeros((ns, n)), soft_weight * np.eye(ns)]])
        f = np.concatenate([f, np.zeros(ns)])
        fth = np.vstack([fth, np.zeros((ns, fth.shape[1]))])
        Aneg = -A0.copy()
        Aneg[ids, n + np.arange(ns)] = -nf                            # the slack relaxes BOTH sides
    else:
        Aneg = -A0
    A = np.vstack([A0, Aneg])
    b = np.concatenate([bu, -bl])
    W = np.vstack([W0, -W0])
    src = [(j, 1) for j in range(m)] + [(j, -1) for j in range(m)]
    keep = b < BIG
    return H, f, fth, A[keep], b[keep], W[keep], [sr for sr, k in zip(src, keep) if k]


def _cheb(G, g):
    """Chebyshev radius / centre of {x: G x <= g}; radius <= 0: empty or lower-dimensional."""
    from scipy.optimize import linprog
    nrm = np.linalg.norm(G, axis=1)
    ok = nrm > 1e-13
    if np.any(~ok & (g < -1e-12)):
        return -1.0, None
    G, g, nrm = G[ok], g[ok], nrm[ok]
    nx = G.shape[1]
    c = np.zeros(nx + 1)
    c[-1] = -1.0
    res = linprog(c, A_ub=np.hstack([G, nrm[:, None]]), b_ub=g, bounds=[(None, None)] * nx + [(0, None)],
                  method="highs")
    if res.status != 0:
        return -1.0, None
    return float(res.x[-1]), res.x[:nx]


def active_sets_by_sampling(gpu, lo, hi, n_samples=2_000_000, seed=0, chunk=4_000_000):
    """Distinct optimal active sets of `gpu` (a GpuMPC) over U(lo, hi): (upper words, lower words,
    count) sorted by count.  One batched implicit solve per chunk."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    seen = {}
    done = 0
    while done < n_samples:
        n = min(chunk, n_samples - done)
        Th = np.asfortranarray(lo[:, None] + (hi - lo)[:, None] * rng.random((len(lo), n)))
        out = gpu.solve(Th, want_iters=False)
        ok = out["exitflag"] == 1
        key = np.hstack([out["active_upper"][ok], out["active_lower"][ok]])
        uniq, cnt = np.unique(key, axis=0, return_counts=True)
        for k, c in zip(uniq, cnt):
            t = tuple(int(v) for v in k)
            seen[t] = seen.get(t, 0) + int(c)
        done += n
    words = len(next(iter(seen))) // 2 if seen else 0
    sets = [(k[:words], k[words:], c) for k, c in seen.items()]
    sets.sort(key=lambda s: -s[2])
    return sets


def regions_from_active_sets(q, sets, lo, hi, nu, tol=1e-8, soft_weight=1e6):
    """Exact polyhedron {theta: G theta <= g} and law u0 = F theta + g0 of every active set."""
    H, f, fth, A, b, W, src = make_singlesided(q, soft_weight)
    row_of = {s: i for i, s in enumerate(src)}
    nth = fth.shape[1]
    Hi = np.linalg.inv(H)
    box_G = np.vstack([np.eye(nth), -np.eye(nth)])
    box_g = np.concatenate([np.asarray(hi, float), -np.asarray(lo, float)])
    regions, laws, law_keys = [], [], {}
    for up, lw, cnt in sets:
        act = []
        for j in range(len(q["bu"])):
            if (up[j >> 6] >> (j & 63)) & 1:
                act.append(row_of[(j, 1)])
            elif (lw[j >> 6] >> (j & 63)) & 1:
                act.append(row_of[(j, -1)])
        inact = [i for i in range(A.shape[0]) if i not in act]
        if act:
            Aa = A[act]
            S = Aa @ Hi @ Aa.T
            if np.linalg.cond(S) > 1e12:
                continue
            Si = np.linalg.inv(S)
            Lth = -Si @ (W[act] + Aa @ Hi @ fth)          # lam(theta) = Lth theta + l0 >= 0
            l0 = -Si @ (b[act] + Aa @ Hi @ f)
            Fz = -Hi @ (fth + Aa.T @ Lth)
            gz = -Hi @ (f + Aa.T @ l0)
            G = np.vstack([-Lth, A[inact] @ Fz - W[inact], box_G])
            g = np.concatenate([l0, b[inact] - A[inact] @ gz, box_g])
        else:
            Fz, gz = -Hi @ fth, -Hi @ f
            G = np.vstack([A @ Fz - W, box_G])
            g = np.concatenate([b - A @ gz, box_g])
        r, centre = _cheb(G, g)
        if r is None or r <= tol:
            continue
        # one id per control law; laws of different active sets that agree to 1e-9 are the SAME law (dual
        # degeneracy: several active sets describe one region, common with soft rows next to hard bounds)
        vec = np.concatenate([Fz[:nu].ravel(), gz[:nu]])
        lid = None
        for li, (Fl, gl) in enumerate(laws):
            ref = np.concatenate([Fl.ravel(), gl])
            if np.max(np.abs(vec - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref))):
                lid = li
                break
        if lid is None:
            lid = len(laws)
            laws.append((Fz[:nu].copy(), gz[:nu].copy()))
        regions.append(dict(G=G, g=g, law=lid, count=cnt, n_facets=len(g) - 2 * nth, centre=centre,
                            radius=r, key=(up, lw)))
    return regions, laws


def _lp_min(cvec, c0, G, g):
    """min c'theta + c0 over {G theta <= g} (None when empty / unbounded)"""
    from scipy.optimize import linprog
    res = linprog(cvec, A_ub=G, b_ub=g, bounds=[(None, None)] * len(cvec), method="highs")
    return None if res.status != 0 else float(res.fun) + c0


def build_tree(regions, laws, nth, samples, exact_top=4, clip=None, tol=1e-9):
    """Binary search tree over region facets -> (halfplanes n_hp x (nth+1), hp_list, jump_list) in the
    layout of lmpc_gpu_create_bst: jump 0 = leaf with law hp_list[i]; else test half-plane hp_list[i]
    and go to i + jump (a'theta <= b) or i + jump + 1.  Facets are ranked by how evenly they split
    the sampled parameters of the cell; the best few are scored exactly (LPs decide which regions
    reach which side), so every leaf holds regions of ONE law.

    clip (nu x 2, [umin umax]): the evaluator clips the law's value to these bounds
    (bst.clipping, src/simulation.jl:163; lmpc_gpu_create_bst's `clipping`).  A cell is then a leaf as
    soon as ONE of its laws reproduces every region of the cell after clipping: a region whose output i
    is the constant bound is covered by a law that stays beyond that bound on the region's part of the
    cell (one LP per region and output).  This is what absorbs the saturated regions."""
    nu = laws[0][0].shape[0] if laws else 0
    sat = []                                     # per law, per output: +1 constant at umax, -1 at umin, 0 otherwise
    for F, g0 in laws:
        row = []
        for i in range(nu):
            v = 0
            if clip is not None and np.all(np.abs(F[i]) < 1e-10):
                v = 1 if abs(g0[i] - clip[i][1]) < 1e-8 else (-1 if abs(g0[i] - clip[i][0]) < 1e-8 else 0)
            row.append(v)
        sat.append(row)

    def covering_law(cG, cg, regs, pts):
        """a law that, clipped, equals every region's law on the cell; None if there is none"""
        ids = sorted({regions[r]["law"] for r in regs}, key=lambda l: sum(1 for v in sat[l] if v != 0))
        if len(ids) <= 1:
            return ids[0] if ids else 0
        if clip is None:
            return None
        for ell in ids:
            Fl, gl = laws[ell]
            if len(pts):                         # cheap rejection on the cell's sample points
                val = np.clip(pts @ Fl.T + gl, [c[0] for c in clip], [c[1] for c in clip])
                bad = False
                for r in regs:
                    lr = regions[r]["law"]
                    if lr == ell:
                        continue
                    inside = np.all(pts @ regions[r]["G"].T <= regions[r]["g"] + 1e-9, axis=1)
                    if inside.any():
                        Fr, gr = laws[lr]
                        if np.max(np.abs(val[inside] - (pts[inside] @ Fr.T + gr))) > 1e-7:
                            bad = True
                            break
                if bad:
                    continue
            ok = True
            for r in regs:
                lr = regions[r]["law"]
                if lr == ell:
                    continue
                Fr, gr = laws[lr]
                G = np.vstack([regions[r]["G"], cG]) if len(cg) else regions[r]["G"]
                g = np.concatenate([regions[r]["g"], cg]) if len(cg) else regions[r]["g"]
                for i in range(nu):
                    if np.allclose(Fr[i], Fl[i], atol=1e-10) and abs(gr[i] - gl[i]) < 1e-9:
                        continue
                    sgn = sat[lr][i]
                    if sgn == 0:
                        ok = False
                        break
                    # the region sits at umax (sgn = +1): need min ell_i >= umax; at umin: max ell_i <= umin
                    v = _lp_min(sgn * Fl[i], sgn * gl[i], G, g)
                    bound = clip[i][1] if sgn > 0 else -clip[i][0]
                    if v is None or v < bound - tol:
                        ok = False
                        break
                if not ok:
                    break
            if ok:
                return ell
        return None

    cand, owner, seen = [], [], {}
    for ri, r in enumerate(regions):
        nf = r["n_facets"]
        for row, rhs in zip(r["G"][:nf], r["g"][:nf]):
            nr = np.linalg.norm(row)
            if nr < 1e-12:
                continue
            a, b = row / nr, rhs / nr
            s = 1.0 if a[np.argmax(np.abs(a) > 1e-9)] > 0 else -1.0
            key = tuple(np.round(np.concatenate([s * a, [s * b]]), 8))
            if key not in seen:
                seen[key] = len(cand)
                cand.append((s * a, s * b))
                owner.append(set())
            owner[seen[key]].add(ri)
    # a cloud of interior points per region (samples that fell into it, its Chebyshev centre, and a short
    # hit-and-run walk for the regions too thin to have caught samples): lets a candidate facet be scored by how
    # it splits the REGIONS of a cell (the classical max(#left, #right) criterion, which is what bounds the
    # depth) with plain dot products, before the few best are checked exactly with LPs
    rng = np.random.default_rng(12345)
    clouds = []
    # law of every sample point (through the region it lies in): lets a facet also be scored by how much it
    # tells about the LAW of a parameter drawn from the sampling distribution (entropy of the two sides), which
    # is what keeps the walk short where the parameters actually fall -- e.g. the 95 % of the invpend range
    # where the first move is saturated end after two or three tests
    labels = -np.ones(len(samples), int)
    for ri, r in enumerate(regions):
        if len(samples):
            ins = np.all(samples @ r["G"].T <= r["g"] + 1e-10, axis=1)
            labels[ins & (labels < 0)] = r["law"]
    for r in regions:
        G, g = r["G"], r["g"]
        pts_r = [r["centre"]]
        if len(samples):
            inside = np.all(samples @ G.T <= g + 1e-10, axis=1)
            pts_r.extend(samples[inside][:48])
        x = np.array(r["centre"], float)
        while len(pts_r) < 24:                         # hit and run from the centre
            dvec = rng.standard_normal(nth)
            dvec /= np.linalg.norm(dvec)
            Gd, slack = G @ dvec, g - G @ x
            with np.errstate(divide="ignore", invalid="ignore"):
                tpos = np.where(Gd > 1e-14, slack / Gd, np.inf).min()
                tneg = np.where(Gd < -1e-14, slack / Gd, -np.inf).max()
            if not np.isfinite(tpos) or not np.isfinite(tneg):
                break
            x = x + (tneg + (tpos - tneg) * rng.uniform(0.05, 0.95)) * dvec
            pts_r.append(x.copy())
        clouds.append(np.array(pts_r))
    hp_used, hp_rows = {}, []

    def hp_index(ci):
        if ci not in hp_used:
            hp_used[ci] = len(hp_rows)
            hp_rows.append(np.concatenate([cand[ci][0], [cand[ci][1]]]))
        return hp_used[ci]

    def reaches(region, cG, cg, a, b, eps=1e-7):
        G = np.vstack([region["G"], cG]) if len(cg) else region["G"]
        g = np.concatenate([region["g"], cg]) if len(cg) else region["g"]
        left = _cheb(np.vstack([G, a]), np.concatenate([g, [b]]))[0]
        right = _cheb(np.vstack([G, -a]), np.concatenate([g, [-b]]))[0]
        return (left is not None and left > eps), (right is not None and right > eps)

    def entropy(lab):
        if len(lab) == 0:
            return 0.0
        _, cnt = np.unique(lab, return_counts=True)
        pr = cnt / cnt.sum()
        return float(-(pr * np.log2(pr)).sum())

    def rec(cG, cg, regs, pts, lab=None, depth=0, used=frozenset()):
        if lab is None:
            lab = labels
        ell = covering_law(cG, cg, regs, pts)
        if ell is not None:
            return ("leaf", ell)
        regset = set(regs)
        # (a facet already on the path to this cell cannot split it again: without this, rounding in the LPs
        # can make a flat slice behind the facet look full-dimensional and the same cut is taken for ever)
        cids = [ci for ci in range(len(cand)) if (owner[ci] & regset) and ci not in used]
        # region-level estimate from the clouds (restricted to the cell), sample balance as the tie-break
        cell_clouds = []
        for r in regs:
            cp = clouds[r]
            if len(cg):
                keep = np.all(cp @ cG.T <= cg + 1e-10, axis=1)
                cp = cp[keep] if keep.any() else cp[:1]
            cell_clouds.append(cp)
        scores = []
        for ci in cids:
            a, b = cand[ci]
            nl = nr = 0
            for cp in cell_clouds:
                side = cp @ a <= b
                nl += bool(side.any())
                nr += bool((~side).any())
            if len(pts):
                sl = int(np.count_nonzero(pts @ a <= b))
                sb = max(sl, len(pts) - sl) / max(len(pts), 1)
            else:
                sb = 0.0
            useless = (nl == len(regs) and nr == len(regs)) or nl == 0 or nr == 0
            scores.append((len(regs) + 1 if useless else max(nl, nr), nl + nr, sb))
        ranked = [cids[i] for i in sorted(range(len(cids)), key=lambda i: scores[i])]
        # where the cell holds enough sample mass of more than one law, the facets that say most about the law
        # go first (expected depth); the region-balanced ranking above (worst-case depth) follows
        top = []
        if depth < 10 and len(pts) >= 64 and len(np.unique(lab)) > 1:      # entropy first only near the root
            h0 = entropy(lab)
            gains = []
            for ci in cids:
                a, b = cand[ci]
                side = pts @ a <= b
                nl_ = int(side.sum())
                if nl_ == 0 or nl_ == len(pts):
                    gains.append(0.0)
                    continue
                gains.append(h0 - (nl_ * entropy(lab[side]) + (len(pts) - nl_) * entropy(lab[~side])) / len(pts))
            order = sorted(range(len(cids)), key=lambda i: -gains[i])
            if gains[order[0]] > 0.02:
                top = [cids[i] for i in order[:max(2, exact_top // 2)] if gains[i] > 0.02]
                ranked = top + [c_ for c_ in ranked if c_ not in top]
        by_entropy = bool(top)
        choice = None
        for pos, ci in enumerate(ranked):
            a, b = cand[ci]
            sides = [reaches(regions[r], cG, cg, a, b) for r in regs]
            Lr = [r for r, sd in zip(regs, sides) if sd[0]]
            Rr = [r for r, sd in zip(regs, sides) if sd[1]]
            if Lr and Rr and (len(Lr) < len(regs) or len(Rr) < len(regs)):
                sc = (pos if by_entropy else 0, max(len(Lr), len(Rr)), len(Lr) + len(Rr))
                if choice is None or sc < choice[0]:
                    choice = (sc, ci, Lr, Rr)
            if choice is not None and (by_entropy or pos >= exact_top - 1):
                break
        if choice is None:
            # No facet separates the regions of this cell: they overlap.  That is legitimate where the active
            # set is not unique (dual degeneracy, common with soft rows next to hard bounds): the regions then
            # carry the same control law up to rounding.  Accept when the laws agree on the cell.
            rc, xc = _cheb(cG, cg) if len(cg) else (1.0, np.zeros(nth))
            probe = [xc] if xc is not None else []
            probe += list(pts[:64]) if len(pts) else []
            for r in regs:
                cp = clouds[r]
                keep = np.all(cp @ cG.T <= cg + 1e-10, axis=1) if len(cg) else np.ones(len(cp), bool)
                probe += list(cp[keep][:4])
            lids = sorted({regions[r]["law"] for r in regs})
            P = np.array(probe)
            vals = [P @ laws[l][0].T + laws[l][1] for l in lids]
            spread = max(np.max(np.abs(v - vals[0])) for v in vals) if len(P) else np.inf
            # (or the cell is a sliver along a common boundary, where continuity makes the laws agree to its width)
            if spread <= 1e-6 * max(1.0, np.max(np.abs(vals[0]))):
                return ("leaf", lids[0])
            raise RuntimeError("explicit builder: no separating facet found (laws differ by %.2e on the cell)" % spread)
        _, ci, Lr, Rr = choice
        a, b = cand[ci]
        lg = np.vstack([cG, a]) if len(cg) else a[None, :]
        rg = np.vstack([cG, -a]) if len(cg) else -a[None, :]
        mask = pts @ a <= b if len(pts) else np.zeros(0, bool)
        return ("node", hp_index(ci),
                rec(lg, np.concatenate([cg, [b]]), Lr, pts[mask] if len(pts) else pts, lab[mask] if len(pts) else lab,
                    depth + 1, used | {ci}),
                rec(rg, np.concatenate([cg, [-b]]), Rr, pts[~mask] if len(pts) else pts, lab[~mask] if len(pts) else lab,
                    depth + 1, used | {ci}))

    tree = rec(np.zeros((0, nth)), np.zeros(0), list(range(len(regions))), samples)
    hp_list, jump_list, queue = [0], [0], [(0, tree)]
    while queue:                       # breadth first: the two children of a node are consecutive
        idx, node = queue.pop(0)
        if node[0] == "leaf":
            hp_list[idx], jump_list[idx] = node[1], 0
        else:
            left = len(hp_list)
            hp_list.extend([0, 0])
            jump_list.extend([0, 0])
            hp_list[idx], jump_list[idx] = node[1], left - idx
            queue.append((left, node[2]))
            queue.append((left + 1, node[3]))
    hp = np.array(hp_rows).reshape(-1, nth + 1) if hp_rows else np.zeros((0, nth + 1))
    return hp, np.array(hp_list, np.int32), np.array(jump_list, np.int32)


def adjacent_active_sets(q, regions, known):
    """Candidate neighbours of the known regions by the classical adjacency rule of mp-QP
    (crossing a facet either drops one active row, lam_j = 0, or adds one inactive row that becomes
    tight): every set one row away from a known full-dimensional region.  Whether a candidate really
    is a full-dimensional region is decided afterwards by its Chebyshev LP, so thin regions that no
    sample hit are still found as long as they touch a region that was."""
    m = len(np.asarray(q["bu"]).ravel())
    bu, bl = np.asarray(q["bu"], float).ravel(), np.asarray(q["bl"], float).ravel()
    new = []

    def offer(up, lw):
        t = tuple(up) + tuple(lw)
        if t not in known:
            known.add(t)
            new.append((tuple(up), tuple(lw), 0))

    for r in regions:
        up0, lw0 = r["key"]
        for j in range(m):
            w, bit = j >> 6, 1 << (j & 63)
            up, lw = list(up0), list(lw0)
            if up[w] & bit:                      # drop an active upper row
                up[w] &= ~bit
                offer(up, lw)
            elif lw[w] & bit:                    # drop an active lower row
                lw[w] &= ~bit
                offer(up, lw)
            else:                                # add the row at either bound it has
                if bu[j] < BIG:
                    up2 = list(up0); up2[w] |= bit
                    offer(up2, lw0)
                if -bl[j] < BIG:
                    lw2 = list(lw0); lw2[w] |= bit
                    offer(up0, lw2)
    return new


def build_explicit(q, gpu, lo, hi, nu, n_samples=2_000_000, seed=0, tree_samples=60000, exact_top=4,
                   explore_rounds=6, clipping=None, soft_weight=1e6):
    """mpQP dict + its GpuMPC -> dict(halfplanes, feedbacks, hp_list, jump_list, regions, laws, sets).
    clipping: nu x 2 [umin umax] the evaluator will clip to (None: leaves reproduce the law exactly)."""
    sets = active_sets_by_sampling(gpu, lo, hi, n_samples, seed)
    regions, laws = regions_from_active_sets(q, sets, lo, hi, nu, soft_weight=soft_weight)
    known = {tuple(s[0]) + tuple(s[1]) for s in sets}
    for _ in range(explore_rounds):            # close the set of regions under facet adjacency
        new = adjacent_active_sets(q, regions, known)
        if not new:
            break
        sets = sets + new
        regions, laws = regions_from_active_sets(q, sets, lo, hi, nu, soft_weight=soft_weight)
    if not regions:
        raise RuntimeError("explicit builder: no full-dimensional critical region found in the range")
    nth = np.asarray(q["f_theta"]).shape[1]
    rng = np.random.default_rng(seed + 1)
    pts = np.asarray(lo, float) + (np.asarray(hi, float) - np.asarray(lo, float)) * rng.random((tree_samples, nth))
    clip = None if clipping is None else np.asarray(clipping, float).reshape(nu, 2)
    hp, hl, jl = build_tree(regions, laws, nth, pts, exact_top, clip=clip)
    fb = np.zeros((len(laws), nu, nth + 1))
    for i, (F, g0) in enumerate(laws):
        fb[i, :, :nth] = F
        fb[i, :, nth] = g0
    return dict(halfplanes=hp, feedbacks=fb, hp_list=hl, jump_list=jl, regions=regions, laws=laws,
                n_candidate_sets=len(sets), clipping=clip, depth=tree_depth(hl, jl))


def tree_depth(hp_list, jump_list):
    """(max depth, number of leaves) of a node-array tree"""
    best, leaves, stack = 0, 0, [(0, 0)]
    while stack:
        i, d = stack.pop()
        if jump_list[i] == 0:
            best, leaves = max(best, d), leaves + 1
        else:
            stack.append((i + jump_list[i], d + 1))
            stack.append((i + jump_list[i] + 1, d + 1))
    return best, leaves


# ---- trees computed by the real package: the arrays ParametricDAQP.codegen emits ------------------------
def write_pdaqp_codegen(path_c, halfplanes, feedbacks, hp_list, jump_list, fname="pdaqp", le_child=1,
                        float_type="double"):
    """Write {fname}.c / {fname}.h in the array format ParametricDAQP.codegen uses (the reference calls it at
    src/codegen.jl:66 and writes its own arrays in the same style, write_float_array src/codegen.jl:220-236):
    `type name[N] = {` newline, one `(c_float)value,` per line, `};`; halfplanes (nth+1) x n_hp column-major,
    feedbacks one (nth+1) x nu column-major block per region, hp_list 0-based, jump_list (0 = leaf), followed
    by the evaluate function.  le_child = 1 writes `id = next_id + (val <= b)` -- the orientation the package
    is believed to use; pass trees in THIS library's orientation (le_child = 0) and they are converted."""
    import os
    hp, fb = np.asarray(halfplanes, float), np.asarray(feedbacks, float)
    hl, jl = np.asarray(hp_list, np.int64).copy(), np.asarray(jump_list, np.int64).copy()
    nth, nu = hp.shape[1] - 1, fb.shape[1]
    if le_child == 1:                      # library order -> package order: exchange the children
        h0, j0 = hl.copy(), jl.copy()
        for i in range(len(j0)):
            if j0[i] == 0:
                continue
            c0, c1 = i + j0[i], i + j0[i] + 1
            hl[c0], jl[c0] = h0[c1], (j0[c1] + 1 if j0[c1] else 0)
            hl[c1], jl[c1] = h0[c0], (j0[c0] - 1 if j0[c0] else 0)
    up = fname.upper()
    base = os.path.splitext(path_c)[0]
    with open(base + ".h", "w") as fh:
        fh.write(f"#ifndef {up}_H\n#define {up}_H\n\ntypedef {float_type} c_float;\ntypedef unsigned short c_int;\n")
        fh.write(f"#define {up}_N_PARAMETER {nth}\n#define {up}_N_SOLUTION {nu}\n\n")
        fh.write(f"extern c_float {fname}_halfplanes[{hp.size}];\nextern c_float {fname}_feedbacks[{fb.size}];\n")
        fh.write(f"extern c_int {fname}_hp_list[{len(hl)}];\nextern c_int {fname}_jump_list[{len(jl)}];\n")
        fh.write(f"void {fname}_evaluate(c_float* parameter, c_float* solution);\n#endif // ifndef {up}_H\n")
    with open(base + ".c", "w") as fc:
        fc.write(f'#include "{os.path.basename(base)}.h"\n')
        for name, arr in ((f"{fname}_halfplanes", hp.ravel()), (f"{fname}_feedbacks", fb.ravel())):
            fc.write(f"c_float {name}[{arr.size}] = {{\n" + "".join(f"(c_float){v:.20f},\n" for v in arr) + "};\n")
        for name, arr in ((f"{fname}_hp_list", hl), (f"{fname}_jump_list", jl)):
            fc.write(f"c_int {name}[{len(arr)}] = {{\n" + "".join(f"(c_int){int(v)},\n" for v in arr) + "};\n")
        branch = "id = next_id + (val <= %s_halfplanes[disp]);" % fname if le_child == 1 else \
                 "id = next_id + (val > %s_halfplanes[disp]);" % fname
        fc.write(f"""void {fname}_evaluate(c_float* parameter, c_float* solution){{
    int i,j,disp;
    int id,next_id;
    c_float val;
    id = 0;
    next_id = id+{fname}_jump_list[id];
    while(next_id != id){{
        // Compute halfplane value
        disp = {fname}_hp_list[id]*({up}_N_PARAMETER+1);
        for(i=0, val=0; i<{up}_N_PARAMETER; i++)
            val += parameter[i] * {fname}_halfplanes[disp++];
        {branch}
        next_id = id+{fname}_jump_list[id];
    }}
    // Leaf node reached -> evaluate affine function
    disp = {fname}_hp_list[id]*({up}_N_PARAMETER+1)*{up}_N_SOLUTION;
    for(i=0; i < {up}_N_SOLUTION; i++){{
        for(j=0, val=0; j < {up}_N_PARAMETER; j++)
            val += parameter[j] * {fname}_feedbacks[disp++];
        val += {fname}_feedbacks[disp++];
        solution[i] = val;
    }}
}}
""")


def load_pdaqp_codegen(path_c, path_h=None):
    """Read the arrays of a ParametricDAQP.codegen output ({fname}.c, optionally {fname}.h for the dimensions).
    Returns dict(halfplanes n_hp x (nth+1), feedbacks n_fb x nu x (nth+1), hp_list, jump_list, nth, nu,
    le_child): le_child is read off the generated evaluate function itself (`next_id + (val <= ...)` or an
    if / else that adds 1 on the <= branch means the <= child is next + 1), so the caller does not have to
    know the package's convention; None when the function is not in the file."""
    import os
    import re
    src = open(path_c).read()
    hdr = open(path_h).read() if path_h and os.path.exists(path_h) else ""
    if not hdr:
        guess = os.path.splitext(path_c)[0] + ".h"
        hdr = open(guess).read() if os.path.exists(guess) else ""
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def array(suffix):
        mt = re.search(r"(\w+)" + suffix + r"\s*\[\s*\d*\s*\]\s*=\s*\{(.*?)\}\s*;", code, flags=re.S)
        if not mt:
            raise ValueError(f"no *{suffix} array in {path_c}")
        body = re.sub(r"\(\s*[\w\s]+\s*\)", "", re.sub(r"//[^\n]*", "", mt.group(2)))   # drop comments, casts
        vals = [v for v in re.split(r"[,\s]+", body) if v]
        return mt.group(1), np.array([float(v.rstrip("fFlLuU")) for v in vals])

    fname, hp = array("_halfplanes")
    _, fb = array("_feedbacks")
    _, hl = array("_hp_list")
    _, jl = array("_jump_list")
    mt = re.search(r"#define\s+\w*N_PARAMETER\s+(\d+)", hdr + src)
    ms = re.search(r"#define\s+\w*N_SOLUTION\s+(\d+)", hdr + src)
    if not mt or not ms:
        raise ValueError("N_PARAMETER / N_SOLUTION not found (pass the .h file)")
    nth, nu = int(mt.group(1)), int(ms.group(1))
    if hp.size % (nth + 1) or fb.size % ((nth + 1) * nu):
        raise ValueError("array sizes do not match N_PARAMETER / N_SOLUTION")
    le_child = None
    body = re.sub(r"//[^\n]*", "", code)
    if re.search(r"next_id\s*\+\s*\(\s*val\s*<=", body) or \
            re.search(r"if\s*\(\s*val\s*<=[^)]*\)\s*\{?\s*id\s*=\s*next_id\s*\+\s*1", body):
        le_child = 1
    elif re.search(r"next_id\s*\+\s*\(\s*val\s*>", body) or \
            re.search(r"if\s*\(\s*val\s*<=[^)]*\)\s*\{?\s*id\s*=\s*next_id\s*;", body):
        le_child = 0
    return dict(halfplanes=hp.reshape(-1, nth + 1), feedbacks=fb.reshape(-1, nu, nth + 1),
                hp_list=hl.astype(np.int32), jump_list=jl.astype(np.int32), nth=nth, nu=nu, le_child=le_child,
                fname=fname)


def verify(empc, gpu, lo, hi, n=1_000_000, seed=123, tol=1e-9):
    """Explicit vs implicit on fresh parameters (the reference asserts 1e-10 at runtests.jl:381):
    returns dict(max_err, frac_within_tol)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    Th = np.asfortranarray(lo[:, None] + (hi - lo)[:, None] * rng.random((len(lo), n)))
    U, _ = empc.evaluate(Th)
    ref = gpu.solve(Th, want_iters=False, want_active=False)
    ok = ref["exitflag"] >= 1
    err = np.max(np.abs(np.asarray(U)[:, ok] - ref["U"][:, ok]), axis=0)
    return dict(max_err=float(err.max()) if err.size else 0.0, frac_within_tol=float(np.mean(err <= tol)),
                n=int(ok.sum()))
