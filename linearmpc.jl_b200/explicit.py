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

Hard constraints only (soft rows change the single-sided form, utils.jl:345-380); offline code,
numpy + scipy's LP, nothing here is on the evaluate path.
"""
from __future__ import annotations

import numpy as np

BIG = 1e20


def make_singlesided(q):
    """A z <= b + W theta from the two-sided mpQP dict of MPC.condense() (utils.jl:329-343).
    Returns (A, b, W, src) where src[i] = (two-sided row, +1 upper / -1 lower)."""
    H = np.asarray(q["H"], float)
    n = H.shape[0]
    bu, bl = np.asarray(q["bu"], float).ravel(), np.asarray(q["bl"], float).ravel()
    m = len(bu)
    Ag = np.asarray(q["A"], float).reshape(-1, n)
    ms = m - Ag.shape[0]
    A0 = np.vstack([np.hstack([np.eye(ms), np.zeros((ms, n - ms))]), Ag])
    W0 = np.asarray(q["W"], float).reshape(m, -1)
    if np.any(np.asarray(q["senses"]) & 8):
        raise NotImplementedError("explicit builder: soft constraints are not supported")
    A = np.vstack([A0, -A0])
    b = np.concatenate([bu, -bl])
    W = np.vstack([W0, -W0])
    src = [(j, 1) for j in range(m)] + [(j, -1) for j in range(m)]
    keep = b < BIG
    return A[keep], b[keep], W[keep], [s for s, k in zip(src, keep) if k]


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


def regions_from_active_sets(q, sets, lo, hi, nu, tol=1e-8):
    """Exact polyhedron {theta: G theta <= g} and law u0 = F theta + g0 of every active set."""
    H = np.asarray(q["H"], float)
    f = np.asarray(q["f"], float).ravel()
    fth = np.asarray(q["f_theta"], float)
    A, b, W, src = make_singlesided(q)
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
        key = tuple(np.round(np.concatenate([Fz[:nu].ravel(), gz[:nu]]), 9))
        if key not in law_keys:
            law_keys[key] = len(laws)
            laws.append((Fz[:nu].copy(), gz[:nu].copy()))
        regions.append(dict(G=G, g=g, law=law_keys[key], count=cnt, n_facets=len(g) - 2 * nth, centre=centre,
                            radius=r, key=(up, lw)))
    return regions, laws


def build_tree(regions, laws, nth, samples, exact_top=4):
    """Binary search tree over region facets -> (halfplanes n_hp x (nth+1), hp_list, jump_list) in the
    layout of lmpc_gpu_create_bst: jump 0 = leaf with law hp_list[i]; else test half-plane hp_list[i]
    and go to i + jump (a'theta <= b) or i + jump + 1.  Facets are ranked by how evenly they split
    the sampled parameters of the cell; the best few are scored exactly (LPs decide which regions
    reach which side), so every leaf holds regions of ONE law."""
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
    hp_used, hp_rows = {}, []

    def hp_index(ci):
        if ci not in hp_used:
            hp_used[ci] = len(hp_rows)
            hp_rows.append(np.concatenate([cand[ci][0], [cand[ci][1]]]))
        return hp_used[ci]

    def reaches(region, cG, cg, a, b, eps=1e-9):
        G = np.vstack([region["G"], cG]) if len(cg) else region["G"]
        g = np.concatenate([region["g"], cg]) if len(cg) else region["g"]
        left = _cheb(np.vstack([G, a]), np.concatenate([g, [b]]))[0]
        right = _cheb(np.vstack([G, -a]), np.concatenate([g, [-b]]))[0]
        return (left is not None and left > eps), (right is not None and right > eps)

    def rec(cG, cg, regs, pts):
        ids = {regions[r]["law"] for r in regs}
        if len(ids) <= 1:
            return ("leaf", ids.pop() if ids else 0)
        regset = set(regs)
        cids = [ci for ci in range(len(cand)) if owner[ci] & regset]
        scores = []
        for ci in cids:
            a, b = cand[ci]
            if len(pts):
                nl = int(np.count_nonzero(pts @ a <= b))
                scores.append(max(nl, len(pts) - nl) if 0 < nl < len(pts) else len(pts) + 1)
            else:
                scores.append(0)
        ranked = [cids[i] for i in np.argsort(scores, kind="stable")]
        choice = None
        for pos, ci in enumerate(ranked):
            a, b = cand[ci]
            sides = [reaches(regions[r], cG, cg, a, b) for r in regs]
            Lr = [r for r, sd in zip(regs, sides) if sd[0]]
            Rr = [r for r, sd in zip(regs, sides) if sd[1]]
            if Lr and Rr and (len(Lr) < len(regs) or len(Rr) < len(regs)):
                sc = (max(len(Lr), len(Rr)), len(Lr) + len(Rr))
                if choice is None or sc < choice[0]:
                    choice = (sc, ci, Lr, Rr)
            if choice is not None and pos >= exact_top - 1:
                break
        if choice is None:
            raise RuntimeError("explicit builder: no separating facet found")
        _, ci, Lr, Rr = choice
        a, b = cand[ci]
        lg = np.vstack([cG, a]) if len(cg) else a[None, :]
        rg = np.vstack([cG, -a]) if len(cg) else -a[None, :]
        mask = pts @ a <= b if len(pts) else np.zeros(0, bool)
        return ("node", hp_index(ci), rec(lg, np.concatenate([cg, [b]]), Lr, pts[mask] if len(pts) else pts),
                rec(rg, np.concatenate([cg, [-b]]), Rr, pts[~mask] if len(pts) else pts))

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


def build_explicit(q, gpu, lo, hi, nu, n_samples=2_000_000, seed=0, tree_samples=20000, exact_top=4,
                   explore_rounds=6):
    """mpQP dict + its GpuMPC -> dict(halfplanes, feedbacks, hp_list, jump_list, regions, laws, sets)."""
    sets = active_sets_by_sampling(gpu, lo, hi, n_samples, seed)
    regions, laws = regions_from_active_sets(q, sets, lo, hi, nu)
    known = {tuple(s[0]) + tuple(s[1]) for s in sets}
    for _ in range(explore_rounds):            # close the set of regions under facet adjacency
        new = adjacent_active_sets(q, regions, known)
        if not new:
            break
        sets = sets + new
        regions, laws = regions_from_active_sets(q, sets, lo, hi, nu)
    if not regions:
        raise RuntimeError("explicit builder: no full-dimensional critical region found in the range")
    nth = np.asarray(q["f_theta"]).shape[1]
    rng = np.random.default_rng(seed + 1)
    pts = np.asarray(lo, float) + (np.asarray(hi, float) - np.asarray(lo, float)) * rng.random((tree_samples, nth))
    hp, hl, jl = build_tree(regions, laws, nth, pts, exact_top)
    fb = np.zeros((len(laws), nu, nth + 1))
    for i, (F, g0) in enumerate(laws):
        fb[i, :, :nth] = F
        fb[i, :, nth] = g0
    return dict(halfplanes=hp, feedbacks=fb, hp_list=hl, jump_list=jl, regions=regions, laws=laws,
                n_candidate_sets=len(sets))


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
