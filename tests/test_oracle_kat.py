"""Known-answer tests that pin the CPU oracle (oracle/lfr_oracle.cc).

The reference ships no tests or golden vectors, and its arithmetic
lives in an un-vendored Ceres, so these analytic properties — derivable from
cost.cc and from the definitions of the losses / line search — are what the
oracle is checked against before it is trusted as the parity reference.
"""
import ctypes as C

import numpy as np

from conftest import get_problem


def interp(oracle, grid, row, col):
    g = np.ascontiguousarray(grid, dtype=np.float64).reshape(18)
    f = np.zeros(2); fr = np.zeros(2); fc = np.zeros(2)
    oracle.lib.lfr_ref_interpolate(g.ctypes.data, row, col, f.ctypes.data, fr.ctypes.data, fc.ctypes.data)
    return f, fr, fc


def test_interpolator_reproduces_grid_samples(oracle):
    """Lagrange property of cost.cc:20,22: value at a grid node = the sample."""
    rng = np.random.default_rng(0)
    G = rng.normal(size=(3, 3, 2))
    for i, r in enumerate((-0.5, 0.0, 0.5)):
        for j, c in enumerate((-0.5, 0.0, 0.5)):
            f, _, _ = interp(oracle, G, r, c)
            np.testing.assert_allclose(f, G[i, j], atol=1e-15)


def test_interpolator_constant_grid(oracle):
    G = np.tile(np.array([0.3, -0.7]), (3, 3, 1))
    for r, c in ((0.1, -0.2), (0.49, 0.3), (-0.33, 0.0)):
        f, fr, fc = interp(oracle, G, r, c)
        np.testing.assert_allclose(f, [0.3, -0.7], atol=1e-15)
        np.testing.assert_allclose(fr, 0, atol=1e-14)
        np.testing.assert_allclose(fc, 0, atol=1e-14)


def test_interpolator_exact_for_biquadratics(oracle):
    """Any field of degree <= 2 per axis is reproduced exactly, with derivatives."""
    rng = np.random.default_rng(1)
    co = rng.normal(size=(2, 3, 3))   # f_k(r,c) = sum a_k[p][q] r^p c^q
    pts = (-0.5, 0.0, 0.5)
    G = np.zeros((3, 3, 2))
    for i, r in enumerate(pts):
        for j, c in enumerate(pts):
            for k in range(2):
                G[i, j, k] = sum(co[k, p, q] * r**p * c**q for p in range(3) for q in range(3))
    for r, c in ((0.17, -0.41), (-0.5, 0.5), (0.0, 0.23)):
        f, fr, fc = interp(oracle, G, r, c)
        for k in range(2):
            ex = sum(co[k, p, q] * r**p * c**q for p in range(3) for q in range(3))
            exr = sum(p * co[k, p, q] * r**(p - 1) * c**q for p in range(1, 3) for q in range(3))
            exc = sum(q * co[k, p, q] * r**p * c**(q - 1) for p in range(3) for q in range(1, 3))
            assert abs(f[k] - ex) < 1e-14 and abs(fr[k] - exr) < 1e-13 and abs(fc[k] - exc) < 1e-13


def test_interpolator_clamp(oracle):
    """Outside +-0.5 the value is the clamped point's value and the derivative
    w.r.t. the clamped coordinate is 0; exactly at +-0.5 it is not (cost.cc:17-18,38-43)."""
    rng = np.random.default_rng(2)
    G = rng.normal(size=(3, 3, 2))
    f_in, fr_in, fc_in = interp(oracle, G, 0.5, 0.2)
    f_out, fr_out, fc_out = interp(oracle, G, 0.9, 0.2)
    np.testing.assert_allclose(f_out, f_in, atol=0)
    np.testing.assert_allclose(fr_out, 0, atol=0)
    np.testing.assert_allclose(fc_out, fc_in, atol=0)
    assert np.abs(fr_in).max() > 1e-3          # row_ == row at exactly 0.5
    f2, fr2, fc2 = interp(oracle, G, -0.7, -3.0)
    np.testing.assert_allclose(fr2, 0, atol=0)
    np.testing.assert_allclose(fc2, 0, atol=0)
    np.testing.assert_allclose(f2, G[0, 0], atol=1e-15)


def test_lagrange_derivative_closed_forms(oracle):
    """L' = {4t-1, -8t, 4t+1} (cost.cc:21,23): a grid that is one basis function."""
    for node, dfun in enumerate((lambda t: 4 * t - 1, lambda t: -8 * t, lambda t: 4 * t + 1)):
        G = np.zeros((3, 3, 2))
        G[node, :, 0] = 1.0           # f_0(r, c) = L_node(r)
        for t in (-0.4, 0.0, 0.3):
            _, fr, fc = interp(oracle, G, t, 0.1)
            assert abs(fr[0] - dfun(t)) < 1e-14 and abs(fc[0]) < 1e-14


def loss(oracle, kind, sim, s, **opts):
    rho = np.zeros(3)
    o = oracle.default_options(**opts)
    oracle.lib.lfr_ref_loss(kind, sim, s, C.byref(o), rho.ctypes.data)
    return rho


def test_cauchy_loss(oracle):
    """sim * CauchyLoss(0.25): rho = b log(1 + s/b), b = 1/16 (solve.cc:111)."""
    b = 0.0625
    for s in (0.0, 1e-3, 0.0625, 2.0):
        rho = loss(oracle, 1, 0.9, s)
        np.testing.assert_allclose(rho, [0.9 * b * np.log1p(s / b), 0.9 / (1 + s / b), -0.9 / b / (1 + s / b) ** 2],
                                   rtol=1e-14, atol=1e-300)


def test_tukey_loss_both_variants(oracle):
    """sim * TukeyLoss(0.0625) (solve.cc:120): Ceres 1.x scaling a^2/6 (rho'(0)=1/2),
    Ceres 2.x a^2/3 (rho'(0)=1); outliers (s > a^2) have rho' = 0."""
    a2 = 0.0625 ** 2
    for variant, scale in ((1, 1.0), (2, 2.0)):
        r0 = loss(oracle, 2, 1.0, 0.0, tukey_variant=variant)
        np.testing.assert_allclose(r0, [0.0, 0.5 * scale, -scale / a2], rtol=1e-15)
        s = 0.4 * a2
        v = 1 - s / a2
        r = loss(oracle, 2, 0.8, s, tukey_variant=variant)
        np.testing.assert_allclose(r, [0.8 * scale * a2 / 6 * (1 - v**3), 0.8 * scale * 0.5 * v * v,
                                       -0.8 * scale * v / a2], rtol=1e-14)
        out = loss(oracle, 2, 0.8, 1.5 * a2, tukey_variant=variant)
        np.testing.assert_allclose(out, [0.8 * scale * a2 / 6, 0, 0], rtol=1e-15)
    # rho' is the derivative of rho (finite difference)
    for kind in (1, 2):
        s, h = 0.3 * a2, 1e-9
        d = (loss(oracle, kind, 1.0, s + h)[0] - loss(oracle, kind, 1.0, s - h)[0]) / (2 * h)
        assert abs(d - loss(oracle, kind, 1.0, s)[1]) < 1e-5


def minimize_interp(oracle, samples, lo, hi, fast=False):
    """fast=False: the oracle's default, literal polynomial.cc; fast=True: the normalised formulation
    the GPU mirrors (tests/test_linesearch_modes.py compares the two)."""
    a = np.array(samples, dtype=np.float64)
    x, v = C.c_double(), C.c_double()
    fn = oracle.lib.lfr_ref_minimize_interpolating_polynomial_fast if fast else \
        oracle.lib.lfr_ref_minimize_interpolating_polynomial
    fn(a.ctypes.data, len(samples), lo, hi, C.byref(x), C.byref(v))
    return x.value, v.value


def test_cubic_interpolation_minimiser(oracle):
    """Two samples with value+gradient of a known cubic: the minimiser inside the
    bracket is the cubic's stationary point (polynomial.cc MinimizePolynomial)."""
    p = np.poly1d([2.0, -3.0, -1.0, 4.0])     # p' = 6x^2 - 6x - 1 -> root at (3+sqrt(15))/6 = 1.1455
    dp = p.deriv()
    s = [[0.0, p(0.0), dp(0.0), 1, 1], [2.0, p(2.0), dp(2.0), 1, 1]]
    x, v = minimize_interp(oracle, s, 0.002, 1.2)
    assert abs(x - (3 + np.sqrt(15)) / 6) < 1e-12 and abs(v - p(x)) < 1e-12
    # bracket excludes the stationary point -> best of {middle, ends}
    x2, _ = minimize_interp(oracle, s, 0.002, 0.9)
    assert x2 == 0.9


def test_quintic_interpolation_and_roots(oracle):
    """Three samples with gradients -> degree-5 interpolant; its derivative's
    roots come from the companion-matrix-equivalent root finder."""
    p = np.poly1d([1.0, -2.0, 0.5, 1.0, -0.3, 0.7])
    dp = p.deriv()
    s = [[x, p(x), dp(x), 1, 1] for x in (0.0, 1.0, 0.45)]
    lo, hi = 0.001, 0.6
    x, v = minimize_interp(oracle, s, lo, hi)
    grid = np.linspace(lo, hi, 200001)
    assert v <= p(grid).min() + 1e-12
    assert abs(v - p(x)) < 1e-12
    out = np.zeros(8)
    co = np.array([1.0, -6.0, 11.0, -6.0])           # (x-1)(x-2)(x-3)
    n = oracle.lib.lfr_ref_polynomial_roots(co.ctypes.data, 4, 0.0, 10.0, out.ctypes.data)
    assert n == 3
    np.testing.assert_allclose(out[:3], [1, 2, 3], rtol=1e-14)
    n = oracle.lib.lfr_ref_polynomial_roots(co.ctypes.data, 4, 1.5, 2.5, out.ctypes.data)
    assert n == 1 and abs(out[0] - 2) < 1e-14
    co = np.array([1.0, 0.0, 0.0, 0.0, 4.0])         # x^4 + 4: no real root
    assert oracle.lib.lfr_ref_polynomial_roots(co.ctypes.data, 5, -10.0, 10.0, out.ctypes.data) == 0
    co = np.poly([0.1, 0.35, 0.36, 0.9])              # quartic with two close roots
    n = oracle.lib.lfr_ref_polynomial_roots(np.ascontiguousarray(co).ctypes.data, 5, 0.0, 1.0, out.ctypes.data)
    assert n == 4
    np.testing.assert_allclose(out[:4], [0.1, 0.35, 0.36, 0.9], rtol=1e-10)


def test_interpolant_minimiser_random_samples(oracle):
    """Property test over random line-search states in Ceres' contraction range
    (current step in [1e-3, 0.6] x previous, bracket [1e-3, 0.6] x current): the
    returned minimiser is a global minimiser over the bracket of the degree-3 /
    degree-5 Hermite interpolant built independently with numpy (Vandermonde solve
    in x, polynomial.cc FindInterpolatingPolynomial), up to the conditioning of that
    solve.  Checked for the fast formulation on every state, and for the literal one on the
    states where Ceres' raw-step Vandermonde system keeps full numerical rank (current step
    >= 0.02; below ~2e-3 Eigen's fullPivLu drops the x^5 column and Ceres' own fit is no longer
    the interpolant — tests/test_linesearch_modes.py)."""
    rng = np.random.default_rng(1234)
    for trial in range(800):
        fast = trial % 2 == 0
        trial //= 2
        three = trial % 2 == 1
        x2 = 10.0 ** rng.uniform(-3, 0)                     # previous step
        x1 = x2 * (rng.uniform(0.02, 0.6) if three else 1.0)  # current step
        f0, g0 = rng.normal(), -abs(rng.normal())           # descent direction at 0
        f1, g1, f2, g2 = rng.normal(size=4) * np.array([1.0, 3.0 / x1, 1.0, 3.0 / x2])
        s = [[0.0, f0, g0, 1, 1], [x1, f1, g1, 1, 1]] + ([[x2, f2, g2, 1, 1]] if three else [])
        lo, hi = 1e-3 * x1, 0.6 * x1
        if not fast and x1 < 0.02:
            continue
        x, v = minimize_interp(oracle, s, lo, hi, fast=fast)
        # independent construction, normalised abscissae for conditioning
        h = max(r[0] for r in s)
        rows, rhs = [], []
        deg = 2 * len(s) - 1
        for (xs, fs, gs, _, _) in s:
            t = xs / h
            rows.append([t ** k for k in range(deg, -1, -1)])
            rhs.append(fs)
            rows.append([k * t ** (k - 1) if k > 0 else 0.0 for k in range(deg, -1, -1)])
            rhs.append(gs * h)
        co = np.linalg.solve(np.array(rows), np.array(rhs))
        p = np.poly1d(co)
        cand = [lo / h, hi / h, 0.5 * (lo + hi) / h] + [r.real for r in p.deriv().roots
                                                       if abs(r.imag) < 1e-9 and lo / h <= r.real <= hi / h]
        best = min(p(c) for c in cand)
        scale = max(1.0, np.abs(co).max())
        assert lo <= x <= hi
        assert abs(v - p(x / h)) <= 1e-9 * scale, (trial, v, p(x / h))
        assert v <= best + 1e-9 * scale, (trial, v, best)


def _two_node_problem(t, sim=0.9):
    """One match between two images, constant flows d12 = t, d21 = -t."""
    from lfr_b200 import MatchSet, build_problem
    t = np.asarray(t, dtype=np.float32)
    d12 = np.tile(t, 9).astype(np.float32)[None]
    d21 = np.tile(-t, 9).astype(np.float32)[None]
    ms = MatchSet(image_names=["a.png", "b.png"], pair_img1=np.array([0]), pair_img2=np.array([1]),
                  pair_fact1=np.ones(1, np.float32), pair_fact2=np.ones(1, np.float32),
                  pair_ptr=np.array([0, 1]), feat1=np.array([3], np.uint32), feat2=np.array([5], np.uint32),
                  sim=np.array([sim], np.float32), disp1=d21, disp2=d12)
    return build_problem(ms)


def test_two_node_track_known_answer(oracle):
    """Equal scores: sort+reverse on (score, idx) makes the higher node index the
    root (solve.cc:567-568).  The single residual is linear in the free node, so
    the trajectory is known in closed form: iteration 1 takes the LM step damped
    by D^2 = diag/radius with radius 1e4, i.e. x = -t * r/(r+1); iteration 2's
    step (|dx| ~ 1e-4 |t|) meets the parameter tolerance and — as in Ceres — is
    NOT applied.  With tight tolerances the solve reaches x = -t."""
    t = np.array([0.125, -0.25])
    p = _two_node_problem(t)
    assert p.is_root.tolist() == [0, 1]
    pos, st = oracle.solve(p, oracle.default_options(n_threads=1))
    r = 1e4
    np.testing.assert_allclose(pos[0], -t * r / (r + 1), rtol=1e-12)
    np.testing.assert_allclose(pos[1], 0, atol=0)
    assert st["iterations"][0] == 2 and st["termination"][0] == 2   # LFR_TERM_PARAMETER_TOL
    tight = oracle.default_options(n_threads=1, function_tolerance=1e-16, parameter_tolerance=1e-14,
                                   gradient_tolerance=1e-14)
    pos_t, st_t = oracle.solve(p, tight)
    np.testing.assert_allclose(pos_t[0], -t, atol=1e-10)
    assert st_t["final_cost"][0] < 1e-20


def test_all_zero_grids_give_zero_displacements(oracle):
    """SKIP_REFINEMENT-style input (compute_match_graph.py:150-152)."""
    from lfr_b200 import build_problem, synth
    ms = synth.generate("cfg1")
    ms.disp1[:] = 0
    ms.disp2[:] = 0
    p = build_problem(ms)
    pos, st = oracle.solve(p)
    assert np.all(pos == 0)
    assert np.all(st["iterations"] == 0)


def test_bounds_are_respected_and_active(oracle):
    """Flows far larger than the +-1 box: the solution sits on the bound."""
    p = _two_node_problem(np.array([3.0, -2.5]))
    pos, _ = oracle.solve(p, oracle.default_options(n_threads=1))
    np.testing.assert_allclose(pos[0], [-1.0, 1.0], atol=1e-12)


def test_trust_region_invariants(oracle):
    """Cost never increases; returned cost matches a re-evaluation at the
    returned point; skipped components are untouched."""
    _, p = get_problem("cfg1")
    pos, st = oracle.solve(p, oracle.default_options(n_threads=2))
    solved = st["termination"] != 0
    assert np.all(st["final_cost"][solved] <= st["initial_cost"][solved] + 1e-15)
    assert st["n_solved"] == int((np.diff(p.comp_ptr.astype(np.int64)) > 1).sum())
    # re-solving from the solution: iteration-0 cost equals the reported final cost
    pos2, st2 = oracle.solve(p, oracle.default_options(n_threads=1), positions=pos)
    np.testing.assert_allclose(st2["initial_cost"][solved], st["final_cost"][solved], rtol=1e-12, atol=1e-18)
    assert np.abs(pos).max() <= 1.0
    roots = p.is_root.astype(bool)
    assert np.all(pos[roots] == 0)


def test_thread_count_does_not_change_results(oracle):
    _, p = get_problem("cfg1")
    a, _ = oracle.solve(p, oracle.default_options(n_threads=1))
    b, _ = oracle.solve(p, oracle.default_options(n_threads=8))
    assert np.array_equal(a, b)
