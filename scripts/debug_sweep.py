import os, sys, types, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from diffsound_b200.dmtet.geometry.dmtet_thickness import DMTetGeometry
d = np.load(os.path.join(ROOT, "tests", "golden", "grid64_tets.npz"))
cases = [(1, 0.2, dict()), (1, 0.2, dict(cheb_degree=24)), (1, 0.9, dict()), (2, 0.2, dict()), (2, 0.2, dict(cheb_degree=24)), (2, 0.9, dict())]
for order, coef, kw in cases:
    FLAGS = types.SimpleNamespace(mode_num=32, order=order, mat="Steel", out_dir="/tmp", without_tensorboard=True)
    geo = DMTetGeometry(64, 1.5, FLAGS, grid=(d["vertices"], d["indices"]))
    geo.apply_sdf(lambda v: 0.6 - v.norm(dim=1))
    obj = geo.getMesh(thickness_coef=torch.tensor(coef))
    for k_, v_ in kw.items():
        setattr(obj, k_, v_)
    t0 = time.time()
    try:
        obj.eigen_decomposition()
        torch.cuda.synchronize()
        s = obj.eig_stats
        print(f"order {order} coef {coef} {kw}: n={obj.deform.pattern.n} slivers={obj._has_slivers()} OK its={s['iterations']} precond={s.get('precond','fp32')} lam0={float(obj.eigenvalues[0]):.6e} t={time.time()-t0:.2f}s", flush=True)
    except Exception as e:
        print(f"order {order} coef {coef} {kw}: n={obj.deform.pattern.n} FAIL t={time.time()-t0:.2f}s {str(e)[:150]}", flush=True)
