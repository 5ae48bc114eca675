"""TEST INFRASTRUCTURE ONLY -- brute-force KKT enumeration, the oracle's independent anchor.

For a strictly convex QP the optimum and (when non-degenerate) the optimal active set are
unique, so enumerating every working set of the LDP
    min 1/2 ||v||^2  s.t.  dl <= Mbar v <= du
and keeping the primal- and dual-feasible one pins both without any active-set logic.
Used by tests/test_oracle_golden.py (anchor (i)).
"""
from itertools import product

import numpy as np


def enumerate_ldp(Mbar, du, dl, soft=None, rho=0.0, tol=1e-9):
    """Return (v, active_upper set, active_lower set, margin) or None if infeasible.

    margin = min over (inactive slacks, active |lambda|): the non-degeneracy measure.
    """
    m, n = Mbar.shape
    soft = np.zeros(m, bool) if soft is None else soft
    best = None
    for assign in product((0, 1, -1), repeat=m):
        idx = [i for i, a in enumerate(assign) if a != 0]
        if len(idx) > n:
            continue
        if idx:
            Mw = Mbar[idx]
            dw = np.array([du[i] if assign[i] > 0 else dl[i] for i in idx])
            G = Mw @ Mw.T + np.diag([rho if soft[i] else 0.0 for i in idx])
            try:
                lam = np.linalg.solve(G, -dw)
            except np.linalg.LinAlgError:
                continue
            v = -Mw.T @ lam
            sgn = np.array([assign[i] for i in idx])
            if np.any(lam * sgn < -tol):
                continue
            lam_margin = np.min(np.abs(lam))
        else:
            v = np.zeros(n)
            lam_margin = np.inf
        mv = Mbar @ v
        inact = [i for i in range(m) if assign[i] == 0]
        su, sl = du[inact] - mv[inact], mv[inact] - dl[inact]
        if inact and (np.any(su < -tol) or np.any(sl < -tol)):
            continue
        slack_margin = min(np.min(su), np.min(sl)) if inact else np.inf
        cand = (v, {i for i in idx if assign[i] > 0}, {i for i in idx if assign[i] < 0},
                min(lam_margin, slack_margin))
        # degenerate problems admit several working sets; keep the smallest
        if best is None or len(idx) < len(best[1]) + len(best[2]):
            best = cand
    return best
