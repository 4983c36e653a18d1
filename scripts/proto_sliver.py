"""CPU prototype: why the two-level preconditioner needs many LOBPCG iterations on marching-tets (short-edge / sliver)
meshes, and what fixes it.  Variants of the coarse solve of the V-cycle on tests/golden/marching_tets.npz meshes:
  exact     sparse LU of the Galerkin P1 operator                     (is the coarse solve the culprit?)
  cheb      Chebyshev on the directly assembled P1 operator           (what the CUDA path does)
  reg:tau   Chebyshev on a P1 operator assembled from element Jacobians whose singular values are clamped to
            >= tau * sigma_max (element-wise regularised geometry: same pattern, bounded element condition)
Not part of the product path."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import modal_oracle as mo
from scripts import proto_pmg as pp

STEEL = pp.STEEL


def p1_stiffness(verts, tets, E, nu, tau=0.0):
    """linear-tet elasticity stiffness; tau > 0: singular values of every element Jacobian clamped to >= tau * sigma_max"""
    mu, la = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    X = verts[tets]                                    # (T, 4, 3)
    J = (X[:, 1:] - X[:, :1]).transpose(0, 2, 1)       # columns = edges
    if tau > 0:
        U, S, Vt = np.linalg.svd(J)
        S = np.maximum(S, tau * S[:, :1])
        sign = np.sign(np.linalg.det(J))
        J = (U * S[:, None, :]) @ Vt
    det = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)                            # rows: d(xi_k)/dx
    gref = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)   # (4, 3) d N_a / d xi
    G = np.einsum("ak,tkd->tad", gref, Jinv)           # (T, 4, 3) grad N_a
    V = det / 6.0
    # K_ab[c,d] = V (la G_ac G_bd + mu G_ad G_bc + mu delta_cd G_a.G_b)
    GG = np.einsum("tac,tbd->tabcd", G, G)
    dots = np.einsum("tac,tbc->tab", G, G)
    Ke = la * GG + mu * GG.transpose(0, 1, 2, 4, 3) + mu * dots[..., None, None] * np.eye(3)
    Ke *= V[:, None, None, None, None]
    rows = (3 * tets[:, :, None, None, None] + np.arange(3)[None, None, None, :, None]) + 0 * tets[:, None, :, None, None] + 0 * np.arange(3)[None, None, None, None, :]
    cols = (3 * tets[:, None, :, None, None] + np.arange(3)[None, None, None, None, :]) + 0 * tets[:, :, None, None, None] + 0 * np.arange(3)[None, None, None, :, None]
    n = 3 * verts.shape[0]
    return sp.csr_matrix((Ke.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n))


class ExactCoarse:
    def __init__(self, Ac):
        self.lu = spla.splu((Ac + 1e-8 * sp.identity(Ac.shape[0]) * abs(Ac.diagonal()).mean()).tocsc())
        self.spmm = 0

    def __call__(self, r, z0=None):
        return self.lu.solve(np.asarray(r, dtype=np.float64))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "g32_sphere"
    variants = sys.argv[2:] or ["exact", "cheb", "reg:0.1"]
    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "marching_tets.npz"))
    v, t = torch.tensor(d[f"{tag}_lcc_verts"]), torch.tensor(d[f"{tag}_lcc_tets"].astype(np.int64))
    pv, pt = mo.promote(v, t, 2)
    K, M = mo.assemble(pv, pt, 2, STEEL[1], STEEL[2], STEEL[0])
    n = K.shape[0]
    P, corners = pp.prolongation(pt, pv.shape[0])
    cmap = -np.ones(pv.shape[0], dtype=np.int64); cmap[corners] = np.arange(corners.size)
    ct = cmap[pt.numpy()[:, [0, 2, 4, 9]]]
    cv = pv.numpy().astype(np.float64)[corners]
    # element quality
    X = cv[ct]; J = (X[:, 1:] - X[:, :1])
    S = np.linalg.svd(J, compute_uv=False)
    print(f"{tag}: n={n} coarse n={3 * corners.size}; element sigma_min/sigma_max: min {np.min(S[:, 2] / S[:, 0]):.2e} "
          f"1% {np.percentile(S[:, 2] / S[:, 0], 1):.2e} median {np.median(S[:, 2] / S[:, 0]):.2e}")
    m, nev = 48, 38
    rng = np.random.default_rng(0)
    X0 = rng.standard_normal((n, m))
    p = pv.numpy().astype(np.float64); p = p - p.mean(0)
    X0[:, :6] = 0
    for c in range(3):
        X0[c::3, c] = 1
    X0[0::3, 3], X0[1::3, 3] = -p[:, 1], p[:, 0]
    X0[1::3, 4], X0[2::3, 4] = -p[:, 2], p[:, 1]
    X0[2::3, 5], X0[0::3, 5] = -p[:, 0], p[:, 2]
    ref = None
    for var in variants:
        parts = var.split(":")
        nu = 3
        cdeg = 32
        dt = np.float32 if "f32" in parts else np.float64
        pp.PMG.coarse_matrix = None
        if parts[0] == "reg":
            pp.PMG.coarse_matrix = p1_stiffness(cv, ct, STEEL[1], STEEL[2], float(parts[1]))
        elif parts[0] == "cheb":
            pp.PMG.coarse_matrix = p1_stiffness(cv, ct, STEEL[1], STEEL[2], 0.0)
        pre = pp.PMG(K, P, nu, 8.0, cdeg, 0.4 * cdeg * cdeg, dt)
        if parts[0] == "exact":
            pre.coarse = ExactCoarse((P.T @ K @ P).tocsr())
        t0 = time.time()
        lam, Xs, it, cols = pp.lobpcg(K, M, X0.copy(), nev, pre, verbose="-v" in parts)
        print(f"{var}: its={it} cols={cols} {time.time() - t0:.1f}s lam[6]={lam[6]:.6e}", flush=True)
        if ref is None:
            ref = lam
        else:
            print("   max rel diff of lam[6:38] vs first:", np.abs(lam[6:nev] / ref[6:nev] - 1).max())


if __name__ == "__main__":
    main()
