"""Time ds_eigh_generalized_f64 (the small Rayleigh-Ritz eigen-solve) on LOBPCG-like pencils; prints sweeps and ms per solve.

    DS_EIGH_JACOBI_ROWS=2 python scripts/bench_eigh.py     # round-1 Jacobi kernel (one row per line position)
    python scripts/bench_eigh.py                           # block kernel (two rows per position, four per warp)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsound_b200 import native  # noqa: E402


def pencil(N, seed):
    rng = np.random.default_rng(seed)
    Q = rng.standard_normal((4 * N, N))
    GM = Q.T @ Q / (4 * N) + 0.1 * np.eye(N)
    ev = np.concatenate([np.abs(rng.standard_normal(6)) * 1e-3, 10 ** rng.uniform(7, 12, N - 6)])
    V = np.linalg.qr(rng.standard_normal((N, N)))[0]
    Lm = np.linalg.cholesky(GM)
    GK = Lm @ (V * ev) @ V.T @ Lm.T
    return (GK + GK.T) / 2, GM


def main():
    out = {"what": "ds_eigh_generalized_f64", "jacobi_rows": os.environ.get("DS_EIGH_JACOBI_ROWS", "4"), "cases": []}
    for N in (48, 96, 144):
        GK, GM = pencil(N, N)
        gk, gm = torch.tensor(GK, device="cuda:0"), torch.tensor(GM, device="cuda:0")
        theta, Cm, info = native.eigh_generalized(gk, gm, 1e5)
        import scipy.linalg
        w = scipy.linalg.eigh(GK, GM, eigvals_only=True)
        err = float(np.abs(theta.cpu().numpy() - np.sort(w)).max() / (np.abs(w).max() + 1e5))
        Cn = Cm.cpu().numpy()
        orth = float(np.abs(Cn.T @ GM @ Cn - np.eye(N)).max())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            native.eigh_generalized(gk, gm, 1e5)
        e1.record()
        torch.cuda.synchronize()
        out["cases"].append({"N": N, "ms": e0.elapsed_time(e1) / reps, "sweeps": int(info.cpu()[1]), "rel_err": err, "orth": orth})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
