"""Powell dog-leg least squares as run by ``ch.minimize(method='dogleg')`` (oracle; test infra only).

The reference calls it at chmosh.py:651-653 (e_3=1e-3), 669-671 and 703-705 (e_3=1e-2), always
with delta_0=0.5 and maxiter=cfg.opt_settings.maxiter.  chumpy itself is not in the tree
(requirements.txt:2, unpinned; 0.70 on PyPI); this restates the published control flow of
``chumpy.optimization_internal._minimize_dogleg`` / ``DoglegState``:
defaults e_1 = e_2 = 1e-15, trust-region update lb=.05 / ub=.9, dense ``np.linalg.solve`` with an
``lstsq`` fallback for the Gauss-Newton step.  PARITY UNPINNED -- chumpy cannot be run here.
"""
from __future__ import annotations

import numpy as np


class DoglegStats:
    def __init__(self):
        self.iterations = 0
        self.r_evals = 0
        self.j_evals = 0
        self.stop_reason = ''


def minimize_dogleg(obj, x0, e_3=0.0, delta_0=None, maxiter=100, e_1=1e-15, e_2=1e-15):
    """obj(x, want_jac) -> r or (r, J).  Returns (x, stats)."""
    st = DoglegStats()
    p = np.asarray(x0, dtype=np.float64).copy()
    r, J = obj(p, True)
    st.r_evals += 1
    st.j_evals += 1
    A = J.T.dot(J)
    g = J.T.dot(-r)
    delta = delta_0
    done = False
    if np.linalg.norm(g, np.inf) < e_1:
        done, st.stop_reason = True, 'small gradient'
    while not done:
        st.iterations += 1
        Jg = J.dot(g)
        d_sd = (g.dot(g) / Jg.dot(Jg)) * g
        d_gn = None
        while True:
            # --- update_step
            if delta is not None and np.linalg.norm(d_sd) >= delta:
                d_dl = (delta / np.linalg.norm(d_sd)) * d_sd
            else:
                if d_gn is None:
                    try:
                        d_gn = np.linalg.solve(A, g)
                    except np.linalg.LinAlgError:
                        d_gn = np.linalg.lstsq(A, g, rcond=None)[0]
                if delta is None or np.linalg.norm(d_gn) <= delta:
                    d_dl = d_gn.copy()
                    if delta is None:
                        delta = np.linalg.norm(d_gn)
                else:
                    delta_sq = delta ** 2
                    diff = d_gn - d_sd
                    sqnorm_sd = d_sd.dot(d_sd)
                    pnow = diff.dot(diff) * delta_sq + d_gn.dot(d_sd) ** 2 - d_gn.dot(d_gn) * sqnorm_sd
                    beta = (delta_sq - sqnorm_sd) / (diff.dot(d_sd) + np.sqrt(pnow))
                    d_dl = d_sd + beta * diff
            improved = False
            if np.linalg.norm(d_dl) <= e_2 * np.linalg.norm(p):
                done, st.stop_reason = True, 'small step'
            else:
                r_trial = obj(p + d_dl, False)
                st.r_evals += 1
                sse0, sse1 = r.dot(r), r_trial.dot(r_trial)
                rho = sse0 - sse1
                if rho > 0:
                    with np.errstate(divide='ignore', invalid='ignore'):
                        rho = rho / (2.0 * g.dot(d_dl) - d_dl.dot(A.dot(d_dl)))
                improved = rho > 0
                if improved:
                    p = p + d_dl
                    if e_3 > 0.0 and (sse0 - sse1) / sse0 < e_3:
                        done, st.stop_reason = True, 'small improvement'
                    else:
                        r_new, J = obj(p, True)
                        st.j_evals += 1
                        r = r_trial
                        A = J.T.dot(J)
                        g = J.T.dot(-r)
                        if np.linalg.norm(g, np.inf) < e_1:
                            done, st.stop_reason = True, 'small gradient'
                # --- updateRadius
                if rho > 0.9:
                    delta = max(delta, 2.5 * np.linalg.norm(d_dl))
                elif rho < 0.05:
                    delta *= 0.25
                if delta <= e_2 * np.linalg.norm(p):
                    done, st.stop_reason = True, 'small trust region'
            if done or improved:
                break
        if not done and st.iterations >= maxiter:
            done, st.stop_reason = True, 'maxiter'
    return p, st
