"""TEST INFRASTRUCTURE ONLY -- explicit-MPC fixture builder.

The reference computes explicit solutions with ParametricDAQP.jl (`mpsolve` + `build_tree`,
call sites /root/reference/src/explicit.jl:39,66), which is not under /root/reference and cannot
run here.  To have something for the batched evaluate kernel to walk, this module builds an
explicit solution itself from the reference's single-sided mpQP
(`make_singlesided`, /root/reference/src/utils.jl:329-391, restated in oracle/mpqp.py):

  1. enumerate candidate optimal active sets, keep those whose critical region
     {theta : lambda(theta) >= 0, A z(theta) <= b + W theta, lo <= theta <= hi} is full-dimensional
     (Chebyshev-centre LP);
  2. per region the affine law for the first move, u0 = F'[theta; 1] (out_inds = 1:nu,
     /root/reference/src/explicit.jl:36), merged by equal law;
  3. a binary search tree over region facets (Tondel-style), in the node-array format of
     include/lmpc_gpu.h: jump_list[i] == 0 -> leaf with law hp_list[i]; otherwise half-plane
     hp_list[i], children at i + jump_list[i] (a'theta <= b) and i + jump_list[i] + 1.

Parity target: explicit == implicit oracle on the same theta (the reference asserts
1e-10 at runtests.jl:381,490,711).
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import combinations, product

import numpy as np
from scipy.optimize import linprog


@dataclass
class Region:
    active: tuple          # indices into the single-sided rows
    G: np.ndarray          # region {theta: G theta <= g}
    g: np.ndarray
    law: int               # id into laws
    Fz: np.ndarray         # z(theta) = Fz theta + gz (all n variables)
    gz: np.ndarray


def _cheb(G, g):
    """Chebyshev radius of {x: Gx <= g} (LP); <= 0 means empty / lower-dimensional."""
    nrm = np.linalg.norm(G, axis=1)
    keep = nrm > 1e-13
    if np.any(~keep & (g < -1e-12)):
        return -1.0, None
    G, g, nrm = G[keep], g[keep], nrm[keep]
    nx = G.shape[1]
    c = np.zeros(nx + 1); c[-1] = -1.0
    res = linprog(c, A_ub=np.hstack([G, nrm[:, None]]), b_ub=g,
                  bounds=[(None, None)] * nx + [(0, None)], method="highs")
    if res.status != 0:
        return -1.0, None
    return res.x[-1], res.x[:nx]


def critical_regions(ss, lo, hi, nu, max_active=None, tol=1e-8):
    """Enumerate full-dimensional critical regions of the single-sided mpQP `ss`."""
    H, f, f_th, A, b, W = ss["H"], ss["f"], ss["f_theta"], ss["A"], ss["b"], ss["W"]
    n, nth, mA = H.shape[0], f_th.shape[1], A.shape[0]
    Hi = np.linalg.inv(H)
    box_G = np.vstack([np.eye(nth), -np.eye(nth)])
    box_g = np.concatenate([hi, -lo])
    max_active = n if max_active is None else max_active
    # pairs (i, i') of rows that are the two sides of the same two-sided row can't both be active
    opposite = {}
    for i in range(mA):
        for j in range(i + 1, mA):
            if np.allclose(A[i], -A[j]) and np.allclose(W[i], -W[j]):
                opposite[i] = j; opposite[j] = i
    regions, laws = [], []

    def law_id(F, g0):
        key = np.round(np.concatenate([F.ravel(), g0]), 9)
        for i, (k, _, _) in enumerate(laws):
            if np.array_equal(k, key):
                return i
        laws.append((key, F, g0))
        return len(laws) - 1

    for ka in range(0, max_active + 1):
        for act in combinations(range(mA), ka):
            if any(opposite.get(i) in act for i in act):
                continue
            act = list(act)
            inact = [i for i in range(mA) if i not in act]
            if ka:
                Aa = A[act]
                S = Aa @ Hi @ Aa.T
                if np.linalg.cond(S) > 1e12:
                    continue
                Si = np.linalg.inv(S)
                # lam = -Si (b + W th + Aa Hi (f + f_th th))
                Lth = -Si @ (W[act] + Aa @ Hi @ f_th)
                l0 = -Si @ (b[act] + Aa @ Hi @ f)
                Fz = -Hi @ (f_th + Aa.T @ Lth)
                gz = -Hi @ (f + Aa.T @ l0)
                G = np.vstack([-Lth, A[inact] @ Fz - W[inact], box_G])
                g = np.concatenate([l0, b[inact] - A[inact] @ gz, box_g])
            else:
                Fz, gz = -Hi @ f_th, -Hi @ f
                G = np.vstack([A @ Fz - W, box_G])
                g = np.concatenate([b - A @ gz, box_g])
            r, _ = _cheb(G, g)
            if r is None or r <= tol:
                continue
            regions.append(Region(tuple(act), G, g, law_id(Fz[:nu], gz[:nu]), Fz, gz))
    return regions, [(F, g0) for _, F, g0 in laws]


def _side(region, cell_G, cell_g, a, b, eps=1e-9):
    """Which side(s) of a'theta <= b the set region ∩ cell reaches: (left, right)."""
    G = np.vstack([region.G, cell_G]) if len(cell_g) else region.G
    g = np.concatenate([region.g, cell_g]) if len(cell_g) else region.g
    left = _cheb(np.vstack([G, a]), np.concatenate([g, [b]]))[0]
    right = _cheb(np.vstack([G, -a]), np.concatenate([g, [-b]]))[0]
    return (left is not None and left > eps), (right is not None and right > eps)


def build_bst(regions, laws, nth, rng=None, samples=None, exact_top=6):
    """Binary search tree over region facets.  Returns (halfplanes n_hp x (nth+1), hp_list, jump_list)."""
    rng = np.random.default_rng(0) if rng is None else rng
    # candidate half-planes: unique normalised facets of all regions (box facets excluded)
    cand, owner = [], []
    seen = {}
    for ri, r in enumerate(regions):
        for row, rhs in zip(r.G[:-2 * nth], r.g[:-2 * nth]):
            nr = np.linalg.norm(row)
            if nr < 1e-12:
                continue
            a, b = row / nr, rhs / nr
            # canonical orientation
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

    def rec(cell_G, cell_g, regs, pts):
        """-> ("leaf", law) or ("node", halfplane id, left subtree, right subtree)"""
        ids = {regions[r].law for r in regs}
        if len(ids) <= 1:
            return ("leaf", ids.pop() if ids else 0)
        # choose the facet (of a region in the cell) that best balances the sampled points
        regset = set(regs)
        cids = [ci for ci in range(len(cand)) if owner[ci] & regset]
        scores = []
        for ci in cids:
            a, b = cand[ci]
            if len(pts):
                nl = int(np.count_nonzero(pts @ a <= b))
                sc = max(nl, len(pts) - nl) if 0 < nl < len(pts) else len(pts) + 1
            else:
                sc = 0
            scores.append(sc)
        ranked = [cids[i] for i in np.argsort(scores, kind="stable")]
        # exact (LP) scoring of the best few sample-ranked facets: minimise the larger child
        choice = None
        for pos, ci in enumerate(ranked):
            a, b = cand[ci]
            sides = [_side(regions[r], cell_G, cell_g, a, b) for r in regs]
            L = [r for r, sd in zip(regs, sides) if sd[0]]
            R = [r for r, sd in zip(regs, sides) if sd[1]]
            if L and R and (len(L) < len(regs) or len(R) < len(regs)):
                sc = (max(len(L), len(R)), len(L) + len(R))
                if choice is None or sc < choice[0]:
                    choice = (sc, ci, L, R)
            if choice is not None and pos >= exact_top - 1:
                break
        if choice is None:
            raise RuntimeError("no separating facet found")
        _, ci, L, R = choice
        a, b = cand[ci]
        lg = np.vstack([cell_G, a]) if len(cell_g) else a[None, :]
        rg = np.vstack([cell_G, -a]) if len(cell_g) else -a[None, :]
        mask = pts @ a <= b if len(pts) else np.zeros(0, bool)
        return ("node", hp_index(ci),
                rec(lg, np.concatenate([cell_g, [b]]), L, pts[mask] if len(pts) else pts),
                rec(rg, np.concatenate([cell_g, [-b]]), R, pts[~mask] if len(pts) else pts))

    pts = samples if samples is not None else np.zeros((0, nth))
    tree = rec(np.zeros((0, nth)), np.zeros(0), list(range(len(regions))), pts)
    # flatten breadth-first: the two children of a node take consecutive indices
    hp_list, jump_list, queue = [0], [0], [(0, tree)]
    while queue:
        idx, node = queue.pop(0)
        if node[0] == "leaf":
            hp_list[idx], jump_list[idx] = node[1], 0
        else:
            left = len(hp_list)
            hp_list.extend([0, 0]); jump_list.extend([0, 0])
            hp_list[idx], jump_list[idx] = node[1], left - idx
            queue.append((left, node[2])); queue.append((left + 1, node[3]))
    return (np.array(hp_rows).reshape(-1, nth + 1), np.array(hp_list, np.int32),
            np.array(jump_list, np.int32))


def explicit_solution(mpc, qp, lo, hi, n_samples=20000, seed=0, exact_top=6):
    """Full fixture: regions, laws, tree and the feedback tensor (n_fb x nu x (nth+1))."""
    from . import mpqp as mq
    ss = mq.make_singlesided(qp, soft_weight=mpc.soft_weight)
    regions, laws = critical_regions(ss, lo, hi, mpc.nu)
    nth = qp.nth
    rng = np.random.default_rng(seed)
    pts = lo + (hi - lo) * rng.random((n_samples, nth))
    hp, hp_list, jump_list = build_bst(regions, laws, nth, rng, pts, exact_top)
    fb = np.zeros((len(laws), mpc.nu, nth + 1))
    for i, (F, g0) in enumerate(laws):
        fb[i, :, :nth] = F
        fb[i, :, nth] = g0
    return dict(regions=regions, laws=laws, halfplanes=hp, hp_list=hp_list, jump_list=jump_list,
                feedbacks=fb)
