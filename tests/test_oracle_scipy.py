"""Independent minimum check of the oracle: scipy's bounded
minimiser on the same robust cost, written separately in numpy, must not find a
meaningfully better point than the oracle's LM, and with tight tolerances the
two must agree on the minimiser."""
import numpy as np
from scipy.optimize import minimize

from conftest import get_problem


def lag(t):
    return np.array([2 * t * (t - .5), -4 * (t - .5) * (t + .5), 2 * t * (t + .5)])


def numpy_cost(p, comp_slot, x_free, free_nodes, pos0):
    """0.5 * sum rho over the kept edges of one component (numpy restatement of A.1)."""
    g = p.graph
    nodes = p.comp_nodes[p.comp_ptr[comp_slot]:p.comp_ptr[comp_slot + 1]].astype(int)
    pos = {int(v): pos0[v].copy() for v in nodes}
    for k, v in enumerate(free_nodes):
        pos[v] = x_free[2 * k:2 * k + 2]
    total = 0.0
    for v in nodes:
        for e in range(int(g.row_ptr[v]), int(g.row_ptr[v + 1])):
            d = int(g.edges["dst"][e])
            if p.track[v] == p.track[d]:
                kind = 1
            elif p.comp[v] == p.comp[d]:
                kind = 2
            else:
                continue
            if p.is_root[v] and p.is_root[d]:
                continue
            D = g.edges["flow"][e].astype(np.float64).reshape(3, 3, 2)
            r_, c_ = np.clip(pos[v], -.5, .5)
            f = np.einsum("i,j,ijk->k", lag(r_), lag(c_), D)
            r = pos[d] - pos[v] - f
            s = float(r @ r)
            sim = float(g.edges["sim"][e])
            if kind == 1:
                total += 0.5 * sim * 0.0625 * np.log1p(s / 0.0625)
            else:
                a2 = 0.0625 ** 2
                total += 0.5 * sim * (a2 / 6 * (1 - (1 - s / a2) ** 3) if s <= a2 else a2 / 6)
    return total


def test_numpy_cost_matches_oracle_cost(oracle):
    _, p = get_problem("cfg1")
    pos, st = oracle.solve(p, oracle.default_options(n_threads=1))
    zero = np.zeros_like(pos)
    for slot in range(0, 60, 7):
        nodes = p.comp_nodes[p.comp_ptr[slot]:p.comp_ptr[slot + 1]].astype(int)
        if len(nodes) <= 1:
            continue
        free = [int(v) for v in nodes if not p.is_root[v]]
        c0 = numpy_cost(p, slot, zero[free].reshape(-1), free, zero)
        c1 = numpy_cost(p, slot, pos[free].reshape(-1), free, pos)
        assert abs(c0 - st["initial_cost"][slot]) <= 1e-12 * max(1, c0)
        assert abs(c1 - st["final_cost"][slot]) <= 1e-12 * max(1, c1)


def test_scipy_finds_no_better_minimum(oracle):
    _, p = get_problem("cfg1")
    pos, st = oracle.solve(p, oracle.default_options(n_threads=1))
    tight = oracle.default_options(n_threads=1, function_tolerance=1e-14, parameter_tolerance=1e-13,
                                   gradient_tolerance=1e-13, max_num_iterations=2000)
    pos_t, st_t = oracle.solve(p, tight)
    checked = 0
    for slot in range(0, 100, 9):
        nodes = p.comp_nodes[p.comp_ptr[slot]:p.comp_ptr[slot + 1]].astype(int)
        if len(nodes) <= 1:
            continue
        free = [int(v) for v in nodes if not p.is_root[v]]
        if not free:
            continue
        fun = lambda x: numpy_cost(p, slot, x, free, pos_t)
        res = minimize(fun, pos_t[free].reshape(-1), method="L-BFGS-B", bounds=[(-1, 1)] * (2 * len(free)),
                       options=dict(ftol=1e-15, gtol=1e-10, maxiter=500))
        c_default = st["final_cost"][slot]
        c_tight = st_t["final_cost"][slot]
        # the default (loose, early-stopping) solve is within 1e-3 relative of the local optimum
        assert c_default - res.fun <= 2e-3 * max(res.fun, 1e-6)
        # the tightly converged oracle is at that optimum
        assert c_tight - res.fun <= 1e-9 * max(res.fun, 1e-6) + 1e-14
        assert np.abs(res.x - pos_t[free].reshape(-1)).max() < 5e-4
        checked += 1
    assert checked >= 5
