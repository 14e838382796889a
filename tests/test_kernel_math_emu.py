"""CPU tier: the small numerical bodies the kernels are built from -- the cyclic Jacobi eigen-solver with its rotation pair as a template
parameter, the 6x6 partial-pivot LU on a register copy, the branch-free eigenvalue ordering -- executed by the host emulation harness
against the oracle's plain-loop versions, bit for bit, on random and on degenerate inputs (ties, zeros, already diagonal, rank deficient)."""
import ctypes as C

import numpy as np

from oracle_api import lib as load_oracle


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _cases3(rng):
    out = []
    for _ in range(300):
        m = rng.normal(size=(3, 3)) * 10.0 ** rng.integers(-6, 3)
        s = m @ m.T
        out.append(s)
    for _ in range(100):   # planar / linear point sets: tiny smallest eigenvalues
        pts = rng.normal(size=(12, 3)) * np.array([1.0, 0.5, 10.0 ** rng.integers(-9, -1)])
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        pts = pts @ q.T + 10.0
        d = pts - pts.mean(0)
        out.append(d.T @ d / 12)
    out += [np.zeros((3, 3)), np.eye(3), np.diag([3.0, 1.0, 2.0]), np.diag([1.0, 1.0, 0.0]), np.ones((3, 3)),
            np.array([[2.0, 1.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.0, 2.0]]), np.array([[1.0, 0.0, 1e-300], [0.0, 1.0, 0.0], [1e-300, 0.0, 1.0]])]
    return out


def test_jacobi_rotation_template_equals_oracle_loop(emu_lib, built):
    L = load_oracle()
    rng = np.random.default_rng(5)
    for s in _cases3(rng):
        a6 = np.array([s[0, 0], s[0, 1], s[0, 2], s[1, 1], s[1, 2], s[2, 2]])
        d1, v1, d2, v2 = np.zeros(3), np.zeros(9), np.zeros(3), np.zeros(9)
        emu_lib.emu_jacobi3(_p(a6), _p(d1), _p(v1))
        L.orc_jacobi_eig3(_p(a6), _p(d2), _p(v2))
        assert np.array_equal(d1, d2) and np.array_equal(v1, v2)


def test_lu6_register_copy_equals_oracle_loop(emu_lib, built):
    L = load_oracle()
    rng = np.random.default_rng(6)
    mats = [np.eye(6) + rng.normal(size=(6, 6)) * 10.0 ** rng.integers(-8, 2) for _ in range(300)]
    mats += [rng.normal(size=(6, 6)) for _ in range(100)]                      # pivoting in every column
    perm = np.eye(6)[[3, 0, 5, 1, 2, 4]]
    mats += [perm, perm * 2.0 + 1e-12, np.diag([1.0, -2.0, 3.0, -4.0, 5.0, -6.0]),
             np.array([[1.0 if abs(i - j) <= 1 else 0.0 for j in range(6)] for i in range(6)]) + np.eye(6)]   # equal pivot candidates
    for A in mats:
        A = np.ascontiguousarray(A)
        o1, o2 = np.zeros(36), np.zeros(36)
        emu_lib.emu_lu_inverse6(_p(A), _p(o1))
        L.orc_lu_inverse6(_p(A), _p(o2))
        assert np.array_equal(o1, o2, equal_nan=True)


def test_order3_equals_bubble_sort_over_indices(emu_lib):
    rng = np.random.default_rng(7)
    vals = [rng.normal(size=3) for _ in range(200)]
    vals += [np.array(v, float) for v in ([1, 1, 1], [1, 1, 0], [0, 1, 1], [1, 0, 1], [2, 1, 1], [1, 2, 1], [1, 1, 2], [3, 2, 1], [0.0, -0.0, 0.0])]
    for ev in vals:
        order = [0, 1, 2]
        for a in range(2):
            for b in range(2 - a):
                if ev[order[b + 1]] < ev[order[b]]:
                    order[b], order[b + 1] = order[b + 1], order[b]
        o = (C.c_int * 3)()
        emu_lib.emu_order3(_p(np.ascontiguousarray(ev)), o)
        assert list(o) == order
