"""CPU prototype (scipy) used to choose the eigensolver's preconditioner: block-Jacobi Chebyshev of
degree d versus a two-level p-multigrid V-cycle (P2 smoother + P1 coarse Chebyshev), optionally with
the preconditioner evaluated in fp32.  Counts LOBPCG iterations / fine-SpMM equivalents to reach the
solver tolerance on quadratic Kuhn cubes.  Not part of the product path."""
import sys, os, time
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import modal_oracle as mo

STEEL = (7850.0, 2.0e11, 0.29)


def block_jacobi_inv(K):
    n = K.shape[0]
    nb = n // 3
    D = np.zeros((nb, 3, 3))
    Kc = K.tocsr()
    for c in range(3):
        for d in range(3):
            D[:, c, d] = np.asarray(Kc[np.arange(nb) * 3 + c, np.arange(nb) * 3 + d]).ravel()
    Di = np.linalg.inv(D)
    rows = (np.arange(nb)[:, None, None] * 3 + np.arange(3)[None, :, None]).repeat(3, 2).ravel()
    cols = (np.arange(nb)[:, None, None] * 3 + np.arange(3)[None, None, :]).repeat(3, 1).ravel()
    return sp.csr_matrix((Di.ravel(), (rows, cols)), shape=(n, n))


def est_lmax(A, Dinv, iters=24, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((A.shape[0], 4))
    for _ in range(iters):
        y = x + Dinv @ (A @ x)
        nx = np.linalg.norm(x, axis=0)
        ny = np.linalg.norm(y, axis=0)
        x = y / ny
    return float((ny / nx).max() - 1.0)


class Cheb:
    """z ~= A^-1 r by `degree` Chebyshev steps on Dinv A over [lmax/ratio, lmax]; x0 optional."""

    def __init__(self, A, Dinv, degree, ratio, dtype=np.float64):
        self.A, self.Dinv = A.astype(dtype), Dinv.astype(dtype)
        self.degree = degree
        lmax = 1.1 * est_lmax(A, Dinv)
        self.lmax, self.lmin = lmax, lmax / ratio
        self.dtype = dtype
        self.spmm = 0

    def __call__(self, r, z0=None):
        r = r.astype(self.dtype)
        theta = 0.5 * (self.lmax + self.lmin)
        delta = 0.5 * (self.lmax - self.lmin)
        sig = theta / delta
        rho = 1.0 / sig
        if z0 is None:
            z = (self.Dinv @ r) / theta
            zp = np.zeros_like(z)
            d = z.copy()
        else:
            z0 = z0.astype(self.dtype)
            res = r - self.A @ z0
            self.spmm += 1
            d = (self.Dinv @ res) / theta
            zp = z0
            z = z0 + d
        for _ in range(1, self.degree):
            rho_new = 1.0 / (2.0 * sig - rho)
            res = r - self.A @ z
            self.spmm += 1
            zn = z + rho_new * rho * (z - zp) + (2.0 * rho_new / delta) * (self.Dinv @ res)
            zp, z = z, zn
            rho = rho_new
        return z


def prolongation(pt, n_nodes):
    """P2 <- P1 interpolation on nodes (scalar), corner list."""
    t = pt.numpy()
    corners = np.unique(t[:, [0, 2, 4, 9]])
    cid = -np.ones(n_nodes, dtype=np.int64)
    cid[corners] = np.arange(corners.size)
    rows, cols, vals = [corners], [cid[corners]], [np.ones(corners.size)]
    for mloc, (a, b) in {1: (0, 2), 3: (2, 4), 5: (4, 0), 6: (0, 9), 7: (2, 9), 8: (4, 9)}.items():
        mid = t[:, mloc]
        for par in (a, b):
            rows.append(mid); cols.append(cid[t[:, par]]); vals.append(np.full(mid.size, 0.5))
    r = np.concatenate(rows); c = np.concatenate(cols); v = np.concatenate(vals)
    key = r * corners.size + c
    _, first = np.unique(key, return_index=True)
    Pn = sp.csr_matrix((v[first], (r[first], c[first])), shape=(n_nodes, corners.size))
    return sp.kron(Pn, sp.identity(3), format="csr"), corners


class PMG:
    """symmetric V(nu,nu): Chebyshev-Jacobi smoother on P2, coarse Chebyshev on P1 (Galerkin)."""

    def __init__(self, A, P, nu, sm_ratio, coarse_degree, coarse_ratio, dtype=np.float64):
        self.A = A.astype(dtype)
        self.P = P.astype(dtype)
        Dinv = block_jacobi_inv(A)
        self.sm = Cheb(A, Dinv, nu, sm_ratio, dtype)
        Ac = (P.T @ A @ P).tocsr() if getattr(PMG, "coarse_matrix", None) is None else PMG.coarse_matrix
        self.nc = Ac.shape[0]
        self.coarse = Cheb(Ac, block_jacobi_inv(Ac), coarse_degree, coarse_ratio, dtype)
        self.dtype = dtype
        self.fine_spmm = 0
        self.ratio_c = Ac.nnz / A.nnz

    def __call__(self, r):
        r = r.astype(self.dtype)
        z = self.sm(r)                         # pre-smooth from zero: nu-1 spmm
        res = r - self.A @ z                   # 1
        zc = self.coarse(self.P.T @ res)
        z = z + self.P @ zc
        z = self.sm(r, z0=z)                   # post-smooth: nu spmm
        return z

    def cost(self):
        return self.sm.spmm + 0  # fine spmm so far (excluding the residual ones, added by caller)


class PMGTwice(PMG):
    """two V-cycles per application: z = V r + V (r - A V r)  (symmetric: 2V - V A V)"""

    def __call__(self, r):
        r = r.astype(self.dtype)
        z = PMG.__call__(self, r)
        return z + PMG.__call__(self, r - self.A @ z)


class PMGDeflated(PMG):
    """V-cycle whose coarse solve is deflated by the lowest coarse eigenvectors Q (from the nested P1
    eigen-solve): zc = Q Th^-1 Q^T rc + Cheb(rc - Mc Q Q^T rc) -- the polynomial then only has to cover the
    spectrum above the deflated modes, so a much lower degree does."""

    def set_deflation(self, Q, theta, MQ, skip=6):
        self.Q = Q[:, skip:].astype(self.dtype)
        self.MQ = MQ[:, skip:].astype(self.dtype)
        self.ith = (1.0 / theta[skip:]).astype(self.dtype)

    def __call__(self, r):
        r = r.astype(self.dtype)
        z = self.sm(r)
        res = r - self.A @ z
        rc = self.P.T @ res
        g = self.Q.T @ rc
        zc = self.Q @ (g * self.ith[:, None]) + self.coarse(rc - self.MQ @ g)
        z = z + self.P @ zc
        z = self.sm(r, z0=z)
        return z


def lobpcg(K, M, X, nev, precond, tol=1e-5, maxit=200, verbose=False):
    n, m = X.shape
    # initial RR
    KX, MX = K @ X, M @ X
    gk, gm = X.T @ KX, X.T @ MX
    th, C = sla.eigh(gk, gm)
    X, KX, MX = X @ C, KX @ C, MX @ C
    lam = th
    P = KP = MP = None
    cols_total = 0
    for it in range(maxit + 1):
        R = KX - MX * lam
        rn = np.linalg.norm(R, axis=0)
        mn = np.linalg.norm(MX, axis=0)
        lref = abs(lam[6])
        scale = np.where(np.arange(m) < 6, lref, np.abs(lam)) * mn
        rel = rn / scale
        act = np.where(rel >= tol)[0]
        nconv = int((rel[:nev] < tol).sum())
        if verbose:
            print(f"  it {it:3d} conv {nconv}/{nev} active {act.size} max rel {rel[6:nev].max():.3e}")
        if nconv >= nev:
            return lam, X, it, cols_total
        W = precond(R[:, act]).astype(np.float64)
        cols_total += act.size
        W -= X @ (MX.T @ W)
        KW, MW = K @ W, M @ W
        if P is not None:
            keep = np.isin(pcols, act)
            S = np.hstack([X, W, P[:, keep]]); KS = np.hstack([KX, KW, KP[:, keep]]); MS = np.hstack([MX, MW, MP[:, keep]])
        else:
            S = np.hstack([X, W]); KS = np.hstack([KX, KW]); MS = np.hstack([MX, MW])
        gk, gm = S.T @ KS, S.T @ MS
        d = 1.0 / np.sqrt(np.diag(gm))
        gk = gk * d[:, None] * d[None, :]; gm = gm * d[:, None] * d[None, :]
        try:
            th, C = sla.eigh(gk, gm)
        except Exception:
            S = np.hstack([X, W]); KS = np.hstack([KX, KW]); MS = np.hstack([MX, MW])
            gk, gm = S.T @ KS, S.T @ MS
            d = 1.0 / np.sqrt(np.diag(gm))
            gk = gk * d[:, None] * d[None, :]; gm = gm * d[:, None] * d[None, :]
            th, C = sla.eigh(gk, gm)
        C = C * d[:, None]
        C = C[:, :m]
        Cp = C.copy(); Cp[:m, :] = 0
        Cp = Cp[:, act]
        P, KP, MP = S @ Cp, KS @ Cp, MS @ Cp
        pcols = act.copy()
        X, KX, MX = S @ C, KS @ C, MS @ C
        lam = th[:m]
    return lam, X, it, cols_total


def main():
    which = sys.argv[2:] or ["cheb", "pmg"]
    if len(sys.argv) > 1 and not sys.argv[1].isdigit():        # a fixture of tests/golden/meshes.npz (bowl, grid16)
        import torch
        d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz"))
        N = sys.argv[1]
        if N.startswith("mtet:"):      # marching-tets output of the reference (sliver elements): tests/golden/marching_tets.npz
            d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "marching_tets.npz"))
            N = N[5:] + "_lcc"
        v, t = torch.tensor(d[f"{N}_verts"]), torch.tensor(d[f"{N}_tets"].astype(np.int64))
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
        v, t = mo.kuhn_cube(N)
    pv, pt = mo.promote(v, t, 2)
    t0 = time.time()
    K, M = mo.assemble(pv, pt, 2, STEEL[1], STEEL[2], STEEL[0])
    n = K.shape[0]
    print(f"N={N} n={n} nnz={K.nnz} assemble {time.time() - t0:.1f}s")
    m, nev = int(os.environ.get("PROTO_M", "48")), int(os.environ.get("PROTO_NEV", "38"))
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
    for w in which:
        if w.startswith("cheb"):
            # cheb[:deg[:f32]]
            parts = w.split(":")
            deg = int(parts[1]) if len(parts) > 1 and parts[1] else int(min(40, max(8, round(n ** (1 / 3) / 3))))
            dt = np.float32 if "f32" in parts else np.float64
            pre = Cheb(K, block_jacobi_inv(K), deg, 0.4 * deg * deg, dt)
            t0 = time.time()
            lam, X, it, cols = lobpcg(K, M, X0.copy(), nev, pre, verbose="-v" in parts)
            print(f"{w}: deg={deg} its={it} fine-spmm-cols={(deg - 1) * cols + 2 * cols}  (per it {deg + 1}) {time.time() - t0:.1f}s")
        elif w.startswith("pmg"):
            # pmg:nu:smratio:cdeg:cratio[:f32]
            parts = w.split(":")
            nu = int(parts[1]) if len(parts) > 1 else 3
            smr = float(parts[2]) if len(parts) > 2 else 8
            cdeg = int(parts[3]) if len(parts) > 3 else 20
            cr = float(parts[4]) if len(parts) > 4 else 0.4 * cdeg * cdeg
            dt = np.float32 if "f32" in parts else np.float64
            P, corners = prolongation(pt, pv.shape[0])
            PMG.coarse_matrix = None
            if "p1" in parts:      # coarse operator = direct P1 assembly on the corner nodes (what the CUDA path does)
                cmap = -np.ones(pv.shape[0], dtype=np.int64); cmap[corners] = np.arange(corners.size)
                import torch as _t
                ct = _t.tensor(cmap[pt.numpy()[:, [0, 2, 4, 9]]])
                K1, _ = mo.assemble(pv[_t.tensor(corners)], ct, 1, STEEL[1], STEEL[2], STEEL[0])
                PMG.coarse_matrix = sp.csr_matrix(K1)
            defl = [q for q in parts if q.startswith("defl")]
            pre = (PMGDeflated if defl else (PMGTwice if "twice" in parts else PMG))(K, P, nu, smr, cdeg, cr, dt)
            t0 = time.time()
            Xs = X0.copy()
            if "nested" in parts:
                Kc = (P.T @ K @ P).tocsr(); Mc = (P.T @ M @ P).tocsr()
                ndeg = int(os.environ.get("PROTO_NESTED_DEG", cdeg))
                cpre = Cheb(Kc, block_jacobi_inv(Kc), ndeg, 0.4 * ndeg * ndeg, dt)
                Xc0 = np.linalg.lstsq((P.T @ P).toarray(), (P.T @ X0)[:, :6], rcond=None)[0] if False else None
                rngc = np.random.default_rng(1)
                Xc = rngc.standard_normal((Kc.shape[0], m))
                Xc[:, :6] = (X0[:, :6])[np.repeat(corners * 3, 3) + np.tile(np.arange(3), corners.size)]
                lamc, Xc, itc, colsc = lobpcg(Kc, Mc, Xc, nev, cpre, tol=float(os.environ.get("PROTO_NESTED_TOL", "3e-2")), verbose="-v" in parts)
                if defl:
                    nd = int(defl[0][4:] or m)
                    pre.set_deflation(Xc[:, :nd], lamc[:nd], Mc @ Xc[:, :nd])
                print(f"   coarse eig: its={itc} coarse-spmm-cols={(cdeg + 1) * colsc} -> fine equiv {(cdeg + 1) * colsc * pre.ratio_c:.0f}; lam err vs fine ref {np.abs(lamc[6:nev] / ref[6:nev] - 1).max() if ref is not None else -1:.3e}")
                Xs = P @ Xc
                Xs[:, :6] = X0[:, :6]
            lam, X, it, cols = lobpcg(K, M, Xs, nev, pre, verbose="-v" in parts)
            per_it = 2 * nu + cdeg * pre.ratio_c + 2
            print(f"{w}: coarse n={pre.nc} nnz ratio {pre.ratio_c:.3f} its={it} fine-spmm-equiv-cols={per_it * cols:.0f} (per it {per_it:.1f}) {time.time() - t0:.1f}s")
        if ref is None:
            ref = lam
        else:
            print("   max rel diff of lam[6:38] vs first:", np.abs(lam[6:nev] / ref[6:nev] - 1).max())


if __name__ == "__main__":
    main()
