"""CPU tier: the pole table of the FP64 solve (csrc/fcn_poles.cu, fcn_common.cuh) against x^(-1/2), and the
whole matrix-function formulation (Householder tridiagonalisation + shifted tridiagonal solves) against the
eigendecomposition the reference uses (module_eigen.f90:37-108), in numpy with the library's own table."""
import ctypes

import numpy as np
import pytest
from scipy.linalg import eigh, hessenberg, solve_banded

from cwbnwp_letkf_b200 import host as H

NP, QMAX = 32, 40


def _table():
    L = H.load_library()
    n = L.letkf_b200_selftest_pole_table(None, 0)
    assert n == (QMAX + 1) * 2 * NP
    t = np.zeros(n)
    assert L.letkf_b200_selftest_pole_table(t.ctypes.data_as(ctypes.c_void_p), n) == n
    return t.reshape(QMAX + 1, 2, NP)


def test_pole_table_approximates_inverse_sqrt():
    t = _table()
    for q in range(1, QMAX + 1):
        c, b = t[q]
        assert (c > 0).all() and (b > 0).all() and (np.diff(b) > 0).all()
        x = np.exp(np.linspace(0.0, q * np.log(2.0), 4001))
        r = (c[None, :] / (x[:, None] + b[None, :])).sum(1)
        err = np.abs(r * np.sqrt(x) - 1.0).max()
        bound = 2e-15 if q <= 20 else (2e-12 if q <= 27 else (5e-10 if q <= 34 else 1e-8))
        assert err < bound, (q, err)
        # slightly outside the interval (rounding of the spectrum bound) the expansion stays accurate
        xo = np.array([1.0 - 1e-6, 2.0 ** q * (1 + 1e-6)])
        ro = (c[None, :] / (xo[:, None] + b[None, :])).sum(1)
        assert np.abs(ro * np.sqrt(xo) - 1.0).max() < max(10 * bound, 1e-12)


@pytest.mark.parametrize("k,p,scale", [(8, 3, 1.0), (32, 5, 1.0), (32, 922, 1.0), (32, 922, 30.0), (96, 1600, 1.0),
                                      (256, 100, 1.0), (256, 922, 100.0)])
def test_matrix_function_form_equals_eigendecomposition_form(k, p, scale):
    """C^(-1/2) x and x . C^-1 b through C = Q T Q^T and the pole expansion, against V diag(f(lambda)) V^T
    (what letkf_solve computes, core:649-668), including rank-deficient Yb (p < k: the eigenvalue mu has
    multiplicity k - p, which an inverse-iteration eigensolver would have to treat as a cluster)."""
    rng = np.random.default_rng(k * 1000 + p)
    t = _table()
    Y = rng.standard_normal((p, k))
    Y -= Y.mean(1, keepdims=True)
    Y *= (np.exp(-0.25 * rng.uniform(0, 13.33, p)) / rng.uniform(0.5, 2.5, p) * scale)[:, None]
    mu = (k - 1) / 1.1
    C = mu * np.eye(k) + Y.T @ Y
    x = rng.standard_normal(k)
    x -= x.mean()
    b = Y.T @ rng.standard_normal(p)
    w, V = eigh(C)
    ref = V @ ((V.T @ x) / np.sqrt(w))
    ref_dot = x @ (V @ ((V.T @ b) / w))
    T, Q = hessenberg(C, calc_q=True)
    d, e = np.diag(T).copy(), np.diag(T, -1).copy()
    lmax = (d + np.abs(np.r_[e, 0]) + np.abs(np.r_[0, e])).max()
    a = mu * (1 - 2.0 ** -20)
    q = max(1, int(np.frexp(lmax / a)[1]))
    c, beta = t[q]

    def rT(z):
        g = np.zeros(k)
        for cj, bj in zip(c, beta):
            ab = np.zeros((3, k))
            ab[1] = d + a * bj
            ab[0, 1:] = e
            ab[2, :-1] = e
            g += cj * np.sqrt(a) * solve_banded((1, 1), ab, z)
        return g

    gx, gb = rT(Q.T @ x), rT(Q.T @ b)
    assert np.abs(Q @ gx - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(gx @ gb - ref_dot) <= 1e-12 * max(abs(ref_dot), np.abs(ref).max() * np.abs(b).max())
