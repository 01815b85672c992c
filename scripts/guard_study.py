"""How well does the growth bound of the static 4 x 4 block pivots predict the rounding error?  (CPU only, numpy.)

Emulates warp_pfaffian_static_steps (csrc/pfaffian_tiles.cuh: natural index order, block pivots, bound s sB / |q|) in
float64 on sums of covariances Lambda_i + Lambda_j (Gram cells, d = 64) of
  * Haar-random pure Gaussian states (no structure), and
  * config C3's kernel ansatz (FermionicPQCKernel, N = 32, depth 2: shallow light cones),
and on Lambda + Lambda_y (projector read-outs of Haar states, random target outcomes), against the oracle's pivoted
Pfaffian.  For every bound: share of the cells handed to the pivoted routine and the largest relative error of the cells
kept on the static path.  Output committed as profiles/r02/guard_study.txt.
    python scripts/guard_study.py [n_samples]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import covariance as cv
from oracle import kernels, pfaffian

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(5)


def rand_so(k):
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def static_pf(a, s_fixed=None):
    """(Pfaffian by static block pivots, largest bound s sB / |q| met on the way)."""
    a = a.copy()
    pf, g, m = 1.0, 0.0, a.shape[0]
    for r0 in range(0, m, 4):
        b = a[r0:r0 + 4, r0:r0 + 4]
        x, y, z, u, v, w = b[0, 1], b[0, 2], b[0, 3], b[1, 2], b[1, 3], b[2, 3]
        q = x * w - y * v + z * u
        sb = abs(x) + abs(y) + abs(z) + abs(u) + abs(v) + abs(w)
        s = s_fixed if s_fixed else np.abs(a[r0:r0 + 4, r0 + 1:]).max()
        if q == 0:
            return None, np.inf
        g = max(g, s * sb / abs(q))
        pf *= q
        if r0 + 4 >= m:
            break
        r = a[r0:r0 + 4, r0 + 4:]
        binv = np.array([[0, -w, v, -u], [w, 0, -z, y], [-v, z, 0, -x], [u, -y, x, 0]]) / q
        a[r0 + 4:, r0 + 4:] += r.T @ (binv @ r)
    return pf, g


def study(name, mats, s_fixed=None):
    ref = np.abs(pfaffian.pfaffian_signed(mats))
    res = []
    for k in range(len(mats)):
        pf, g = static_pf(mats[k], s_fixed)
        res.append((g, abs(abs(pf) - ref[k]) / ref[k] if pf is not None and ref[k] > 0 else np.nan))
    res = np.array(res)
    print(f"{name}: {len(mats)} matrices")
    for lim in (1e5, 3e4, 1e4, 3e3, 1e3):
        keep = res[:, 0] <= lim
        err = res[keep, 1]
        err = err[~np.isnan(err)]
        print(f"  bound <= {lim:7g}: handed over {100 * (1 - keep.mean()):5.1f} %   max rel err kept {err.max():.2e}   "
              f"kept cells with err > 1e-10: {100 * np.mean(err > 1e-10):.2f} %")


d = 64
j0 = np.kron(np.eye(d // 2), np.array([[0.0, 1.0], [-1.0, 0.0]]))
haar = np.stack([q @ j0 @ q.T for q in (rand_so(d) for _ in range(n))])
haar = (haar - haar.transpose(0, 2, 1)) / 2
iu = np.triu_indices(n, 1)
study("Gram cells, Haar-random Gaussian states (d = 64)", (haar[:, None] + haar[None, :])[iu])
x = np.random.default_rng(0).random((n, 64))
k = kernels.FermionicPQCKernelOracle(n_qubits=32, rotations="X,Z").fit(x)
lam0 = cv.covariance_matrix(cv.amplitudes_from_basis_state(np.zeros(32, int)))
c3 = cv.evolve(lam0, k.x_to_sptm(x))
study("Gram cells, config C3 ansatz (N = 32, depth 2, d = 64)", (c3[:, None] + c3[None, :])[iu])
mats = []
for i in range(min(20 * n, 600)):
    lam = haar[i % n]
    y = rng.integers(0, 2, d // 2) * 2 - 1
    mats.append(lam + np.kron(np.diag(y.astype(float)), np.array([[0.0, 1.0], [-1.0, 0.0]])))
study("projector read-outs Lambda + Lambda_y, Haar states, random outcomes (half of them parity forbidden: Pf = 0)",
      np.stack(mats), s_fixed=2.0)
