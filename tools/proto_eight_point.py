"""Prototype (CPU, NumPy) of the K4 arithmetic: Householder null vector + cofactor power
iteration for the rank-2 projection, compared with the reference SVD formulation."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import ransac_oracle as ro

def null_vec_householder(A):           # A: (H, 8, 9) -> (H, 9)
    B = np.transpose(A, (0, 2, 1)).copy()   # (H, 9, 8): columns = design rows
    H = len(B)
    vs, betas = [], []
    for k in range(8):
        x = B[:, k:, k].copy()
        s = np.sum(x * x, axis=1)
        nrm = np.sqrt(s)
        x0 = x[:, 0]
        alpha = np.where(x0 >= 0, -nrm, nrm)
        v = x.copy(); v[:, 0] = x0 - alpha
        beta = np.where(s > 0, 1.0 / (s + np.abs(x0) * nrm), 0.0)
        for j in range(k + 1, 8):
            d = np.sum(v * B[:, k:, j], axis=1) * beta
            B[:, k:, j] -= d[:, None] * v
        vs.append(v); betas.append(beta)
    n = np.zeros((H, 9)); n[:, 8] = 1.0
    for k in range(7, -1, -1):
        d = np.sum(vs[k] * n[:, k:], axis=1) * betas[k]
        n[:, k:] -= d[:, None] * vs[k]
    return n

def cof(F):  # (H,3,3) cofactor matrix
    C = np.empty_like(F)
    for i in range(3):
        for j in range(3):
            i1, i2 = (i + 1) % 3, (i + 2) % 3
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            C[:, i, j] = F[:, i1, j1] * F[:, i2, j2] - F[:, i1, j2] * F[:, i2, j1]
    return C

def v3_power(F, squarings=8):
    C = cof(F)
    M = np.einsum('hki,hkj->hij', C, C)          # C^T C, dominant eigvec = v3
    its = np.zeros(len(F), int)
    for it in range(squarings):
        tr = M[:, 0, 0] + M[:, 1, 1] + M[:, 2, 2]
        M = M / tr[:, None, None]
        M = np.einsum('hik,hkj->hij', M, M)
    d = np.stack([M[:, 0, 0], M[:, 1, 1], M[:, 2, 2]], 1)
    c = np.argmax(d, 1)
    v = M[np.arange(len(F)), :, c]
    return v / np.linalg.norm(v, axis=1, keepdims=True)

def solve(A, squarings=8):
    f = null_vec_householder(A)
    F = f.reshape(-1, 3, 3)
    v3 = v3_power(F, squarings)
    Fv = np.einsum('hij,hj->hi', F, v3)
    return F - Fv[:, :, None] * v3[:, None, :]

def ref(A):
    out = []
    for a in A:
        _, _, Vt = np.linalg.svd(a)
        F = Vt[-1].reshape(3, 3)
        U, S, Vt = np.linalg.svd(F); S[2] = 0
        out.append(U @ np.diag(S) @ Vt)
    return np.stack(out)

def design(src, dst):
    x, y, u, v = src[..., 0], src[..., 1], dst[..., 0], dst[..., 1]
    return np.stack([u * x, u * y, u, v * x, v * y, v, x, y, np.ones_like(x)], -1)

rng = np.random.default_rng(1)
n = 600
P = rng.uniform(-1, 1, (n, 3)) * [10, 2, 17] + [0, 0, 22]
p1 = P[:, :2] / P[:, 2:]
yaw = 0.02
R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
P2 = P @ R.T + [0.05, 0, -1.0]
p2 = P2[:, :2] / P2[:, 2:] + rng.normal(0, 0.5 / 718, (n, 2))
out = rng.permutation(n)[: n // 3]
p2[out] = rng.uniform(-0.8, 0.8, (len(out), 2))
p1 = p1.astype(np.float32).astype(np.float64); p2 = p2.astype(np.float32).astype(np.float64)
Hn = 4000
smp = np.stack([rng.choice(n, 8, replace=False) for _ in range(Hn)])
A = design(p1[smp], p2[smp])
Fr = ref(A)
for sq in (6, 8, 10, 12):
    Fm = solve(A, sq)
    # compare up to sign
    s = np.sign(np.sum(Fr * Fm, axis=(1, 2)))
    err = np.abs(Fr - s[:, None, None] * Fm).max(axis=(1, 2))
    sv = np.linalg.svd(null_vec_householder(A).reshape(-1, 3, 3), compute_uv=False)
    print(sq, "max err", err.max(), "p99", np.quantile(err, .99), "frac>1e-9", (err > 1e-9).mean(), "frac>1e-7", (err > 1e-7).mean(),
          "worst ratio s3/s2", (sv[:, 2] / sv[:, 1])[np.argmax(err)])
# inlier-count parity
th = 0.01
_, c_ref = ro.score_hypotheses(Fr, p1, p2, th)
for sq in (6, 8, 10):
    _, c_new = ro.score_hypotheses(solve(A, sq), p1, p2, th)
    print(sq, "count diffs", np.abs(c_ref - c_new).max(), (c_ref != c_new).sum(), "of", Hn)
