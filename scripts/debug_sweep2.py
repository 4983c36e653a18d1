import os, sys, types, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from diffsound_b200.dmtet.geometry.dmtet_thickness import DMTetGeometry
d = np.load(os.path.join(ROOT, "tests", "golden", "grid64_tets.npz"))
order, coef = int(sys.argv[1]), float(sys.argv[2])
FLAGS = types.SimpleNamespace(mode_num=32, order=order, mat="Steel", out_dir="/tmp", without_tensorboard=True)
geo = DMTetGeometry(64, 1.5, FLAGS, grid=(d["vertices"], d["indices"]))
geo.apply_sdf(lambda v: 0.6 - v.norm(dim=1))
obj = geo.getMesh(thickness_coef=torch.tensor(coef))
obj.eig_maxit = int(sys.argv[3])
obj.cheb_degree = 24
try:
    obj.eigen_decomposition()
    print("OK", obj.eig_stats)
except Exception as e:
    print("FAIL", str(e)[:200])
