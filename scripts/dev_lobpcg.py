"""Dev experiment (GPU): LOBPCG convergence on Kuhn cubes vs scipy eigsh."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DIFFSOUND_PARTIAL_LIB", "1")
from diffsound_b200 import native
from diffsound_b200.diffelastic import mass_matrix as mmx
from oracle import modal_oracle as mo

STEEL = (7850.0, 2.0e11, 0.29)
dev = torch.device("cuda:0")


def run(N, order, k=32, deg=8, ratio=30.0, tol=1e-4, check=False, m=48, maxit=300):
    v, t = mo.kuhn_cube(N)
    t0 = time.time()
    pv, pt = mo.promote(v, t, order)
    verts = pv.to(dev).contiguous(); tets = pt.to(torch.int32).to(dev).contiguous()
    torch.cuda.synchronize(); t1 = time.time()
    pat = native.Pattern(tets, verts.shape[0])
    mu, lam = mo.lame(STEEL[1], STEEL[2])
    ctab = mmx.stiffness_contraction_table(order).to(dev); mtab = mmx.mass_density_table(order, STEEL[0]).to(dev)
    torch.cuda.synchronize(); t2 = time.time()
    Kval, Mblk = native.assemble_km(verts, tets, order, pat, mu, lam, ctab, mtab)
    torch.cuda.synchronize(); t3 = time.time()
    n = pat.n
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g)
    # rigid modes in the first 6 columns
    p = verts.double()
    X[:, :6] = 0
    for c in range(3):
        X[c::3, c] = 1
    X[0::3, 3] = -p[:, 1]; X[1::3, 3] = p[:, 0]
    X[1::3, 4] = -p[:, 2]; X[2::3, 4] = p[:, 1]
    X[2::3, 5] = -p[:, 0]; X[0::3, 5] = p[:, 2]
    torch.cuda.synchronize(); t4 = time.time()
    lamv, res, st = native.lobpcg(pat, Kval, Mblk, X, nev=k + 6, tol=tol, maxit=maxit, cheb_degree=deg, cheb_ratio=ratio,
                                  n_rigid=6, verbose=int(os.environ.get("V", "0")))
    torch.cuda.synchronize(); t5 = time.time()
    print(f"N={N} ord={order} n={n} nnzb={pat.nnzb} promote {t1-t0:.2f}s pattern {t2-t1:.3f}s assemble {t3-t2:.3f}s "
          f"lobpcg {t5-t4:.3f}s deg={deg} ratio={ratio} {st}", flush=True)
    lam_h = lamv.cpu().numpy()
    print("  lam[:10]", lam_h[:10])
    if check:
        K, M = mo.assemble(pv, pt, order, STEEL[1], STEEL[2], STEEL[0])
        t6 = time.time()
        le, U, Uf, S = mo.eig_arpack(K, M, k)
        print(f"  arpack {time.time()-t6:.2f}s  max rel err {np.abs(lam_h[6:6+k]-le).max()/1:.3e} rel "
              f"{(np.abs(lam_h[6:6+k]-le)/le).max():.3e}")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(4, 1, k=16, check=True, m=32)
        run(4, 2, k=16, check=True, m=32)
        run(8, 2, k=32, check=True)
    elif which == "sweep":
        for deg, ratio in ((4, 10.0), (8, 30.0), (12, 60.0), (16, 100.0), (24, 200.0)):
            run(16, 2, deg=deg, ratio=ratio)
        for deg, ratio in ((8, 30.0), (16, 100.0), (24, 200.0), (32, 400.0)):
            run(32, 2, deg=deg, ratio=ratio)
    else:
        run(int(sys.argv[1]), int(sys.argv[2]), deg=int(sys.argv[3]), ratio=float(sys.argv[4]))
