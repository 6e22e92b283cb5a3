"""CPU oracle — 5-point essential-matrix minimal solver.  TEST INFRASTRUCTURE ONLY.

The reference's own RANSAC (homography.py:302-345) is 8-point; its 5-point users are the
``cv2.findEssentialMat`` call sites (slam_viewer.py:195, web_dashboard_server.py:145,
visual_slam_offline_entry_point.py:51), i.e. third-party arithmetic (OpenCV five-point.cpp, Nister
2004 / Stewenius 2006).  This module restates the published algorithm in the form the device
kernel uses (Nister: null space -> ten cubic constraints -> Gauss-Jordan -> 3x3 polynomial
matrix B(z) -> degree-10 polynomial -> real roots -> x, y from B(z)), and is pinned against
cv2.findEssentialMat on exactly five points, which returns every real solution
(tests/golden/fivepoint_golden.npz).  float64.
"""
from __future__ import annotations

import itertools

import numpy as np

# monomials of degree <= 3 in (x, y, z) as exponent triples
_MONO = [m for d in range(4) for m in itertools.product(range(d + 1), repeat=3) if sum(m) == d]
_IDX = {m: i for i, m in enumerate(_MONO)}
# Nister's column order: the ten eliminated monomials first, then x(z^2, z, 1), y(z^2, z, 1), (z^3, z^2, z, 1)
ORDER = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]


def _pmul(a, b):
    """product of two polynomials given as coefficient vectors over _MONO (degree must stay <= 3)"""
    out = np.zeros(len(_MONO))
    for i, ai in enumerate(a):
        if ai == 0.0:
            continue
        for j, bj in enumerate(b):
            if bj == 0.0:
                continue
            m = tuple(p + q for p, q in zip(_MONO[i], _MONO[j]))
            out[_IDX[m]] += ai * bj
    return out


def nullspace_basis(src, dst) -> np.ndarray:
    """(4, 3, 3): E = x E[0] + y E[1] + z E[2] + E[3] spans the solutions of x2^T E x1 = 0."""
    src, dst = np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)
    x, y, u, v = src[:, 0], src[:, 1], dst[:, 0], dst[:, 1]
    Q = np.stack([u * x, u * y, u, v * x, v * y, v, x, y, np.ones_like(x)], axis=1)
    return np.linalg.svd(Q)[2][5:].reshape(4, 3, 3)


def constraint_matrix(basis) -> np.ndarray:
    """The ten cubic constraints (det E = 0, 2 E E^T E - tr(E E^T) E = 0) as a 10x20 matrix in ORDER."""
    lin = np.zeros((3, 3, len(_MONO)))
    for k, m in enumerate([(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]):
        lin[:, :, _IDX[m]] = basis[k]
    E = [[lin[i, j] for j in range(3)] for i in range(3)]
    EEt = [[sum(_pmul(E[i][k], E[j][k]) for k in range(3)) for j in range(3)] for i in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    rows = []
    for i in range(3):
        for j in range(3):
            rows.append(2.0 * sum(_pmul(EEt[i][k], E[k][j]) for k in range(3)) - _pmul(tr, E[i][j]))
    det = (_pmul(_pmul(E[0][0], E[1][1]) - _pmul(E[0][1], E[1][0]), E[2][2])
           - _pmul(_pmul(E[0][0], E[1][2]) - _pmul(E[0][2], E[1][0]), E[2][1])
           + _pmul(_pmul(E[0][1], E[1][2]) - _pmul(E[0][2], E[1][1]), E[2][0]))
    rows.append(det)
    A = np.stack(rows)
    return A[:, [_IDX[m] for m in ORDER]]


def _z_poly(row_a, row_b):
    """<a> - z <b> for two Gauss-Jordan rows (the last ten columns): -> (px[4], py[4], pc[5]), highest power first."""
    a, b = row_a, row_b
    px = np.array([-b[0], a[0] - b[1], a[1] - b[2], a[2]])
    py = np.array([-b[3], a[3] - b[4], a[4] - b[5], a[5]])
    pc = np.array([-b[6], a[6] - b[7], a[7] - b[8], a[8] - b[9], a[9]])
    return px, py, pc


def five_point(src, dst) -> np.ndarray:
    """All real essential matrices through five correspondences -> (k, 3, 3), unit Frobenius norm."""
    basis = nullspace_basis(src, dst)
    A = constraint_matrix(basis)
    G = np.linalg.solve(A[:, :10], A[:, 10:])            # Gauss-Jordan: [I | G]
    rows = [_z_poly(G[4], G[5]), _z_poly(G[6], G[7]), _z_poly(G[8], G[9])]   # (x^2 z, x^2), (y^2 z, y^2), (xyz, xy)
    B = [[np.poly1d(r[0]), np.poly1d(r[1]), np.poly1d(r[2])] for r in rows]
    det = (B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0])
           + B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]))
    sols = []
    for z in np.roots(det.coeffs):
        if abs(z.imag) > 1e-9 * max(1.0, abs(z.real)):
            continue
        z = float(z.real)
        Bz = np.array([[B[i][j](z) for j in range(3)] for i in range(3)])
        # [x, y, 1] spans the null space of B(z): cross product of the two best-conditioned rows
        cands = [np.cross(Bz[0], Bz[1]), np.cross(Bz[0], Bz[2]), np.cross(Bz[1], Bz[2])]
        n = max(cands, key=lambda c: abs(c[2]))
        if n[2] == 0.0:
            continue
        x, y = n[0] / n[2], n[1] / n[2]
        E = x * basis[0] + y * basis[1] + z * basis[2] + basis[3]
        sols.append(E / np.linalg.norm(E))
    return np.stack(sols) if sols else np.zeros((0, 3, 3))


def match_solution_sets(A, B, tol=1e-6):
    """number of matrices of A (unit norm, sign-free) that have a partner in B within tol"""
    hit = 0
    for a in A:
        a = a / np.linalg.norm(a)
        for b in B:
            b = b / np.linalg.norm(b)
            if min(np.abs(a - b).max(), np.abs(a + b).max()) < tol:
                hit += 1
                break
    return hit
