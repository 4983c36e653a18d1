"""Marching tetrahedra -> tet mesh, vertex compaction and largest connected component on the device (csrc/mtet.cu,
diffsound_b200/dmtet/geometry/dmtet_thickness.py) against the unmodified reference run on the CPU
(oracle/make_goldens_r2.py `mtet`: dmtet_thickness.py:99-200, :254-299).

Bar: tets, faces, vertex numbering bit-exact (they define the sparsity pattern downstream); vertex positions bit-exact
(the interpolation is the reference's fp32 operation sequence); d/d(thickness coefficient) <= 1e-5 relative; the
tick() eigen loss <= 1e-5 relative and its gradient to the 32 thickness logits <= 1e-4 rel-L2 (eigenvalue gradient bar)."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _grid(res):
    if res == 16:
        d = np.load(os.path.join(GOLDEN, "meshes.npz"))
        return d["grid16_verts"], d["grid16_tets"].astype(np.int64)
    if res == 32:
        d = np.load(os.path.join(GOLDEN, "mesh_grid32.npz"))
        return d["verts"], d["tets"].astype(np.int64)
    d = np.load(os.path.join(GOLDEN, "grid64_tets.npz"))
    return d["vertices"].astype(np.float32), d["indices"].astype(np.int64)


def _geometry(res, scale, sdf, mode_num=8, order=1):
    from diffsound_b200.dmtet.geometry.dmtet_thickness import DMTetGeometry
    FLAGS = types.SimpleNamespace(mode_num=mode_num, order=order, mat="Steel", out_dir="/tmp", without_tensorboard=True)
    geo = DMTetGeometry(res, scale, FLAGS, grid=_grid(res))
    geo.apply_sdf(torch.tensor(sdf))
    return geo


@pytest.mark.parametrize("tag", ["g16_sphere", "g32_sphere", "g32_two", "g32_full"])
def test_marching_tets_matches_reference(tag):
    g = golden("marching_tets")
    res, scale = int(g[f"{tag}_res"]), float(g[f"{tag}_scale"])
    geo = _geometry(res, scale, g[f"{tag}_sdf"])
    assert abs(float(geo.marching_tets.max_thickness) - float(g[f"{tag}_max_thickness"])) == 0.0
    tc = torch.tensor(float(g[f"{tag}_coef"]), requires_grad=True)
    verts, faces, av, at = geo.marching_tets(geo.verts, geo.sdf, geo.indices, tc)
    assert faces.dtype == torch.int64 and at.dtype == torch.int64
    assert np.array_equal(faces.cpu().numpy(), g[f"{tag}_faces"])
    assert np.array_equal(at.cpu().numpy(), g[f"{tag}_all_tets"])
    assert np.array_equal(verts.detach().cpu().numpy(), g[f"{tag}_verts"])
    assert np.array_equal(av.detach().cpu().numpy(), g[f"{tag}_all_verts"])
    lv, lt = geo.get_largest_connected_component(av, at)
    assert np.array_equal(lt.cpu().numpy(), g[f"{tag}_lcc_tets"])
    assert np.array_equal(lv.detach().cpu().numpy(), g[f"{tag}_lcc_verts"])
    assert (geo.last_n_components > 1) == (g[f"{tag}_lcc_tets"].shape != g[f"{tag}_all_tets"].shape)
    wsum = torch.linspace(0.5, 1.5, lv.shape[0]).unsqueeze(1).to(DEV)
    (lv * wsum).sum().backward()
    ref = float(g[f"{tag}_grad_coef"])
    assert abs(float(tc.grad) - ref) <= 1e-5 * max(abs(ref), 1.0)


def test_components_against_scipy_on_random_forest():
    """union-find labels = smallest vertex id of the component; sizes and the largest-component rule against SciPy."""
    import scipy.sparse as sp
    import scipy.sparse.csgraph as csg
    from diffsound_b200 import native
    rng = np.random.default_rng(7)
    n = 5000
    tets = np.concatenate([rng.integers(0, 1200, (700, 4)), 1200 + rng.integers(0, 800, (300, 4)),
                           2000 + rng.integers(0, 3000, (900, 4))]).astype(np.int64)
    ncomp, kept, tout, labels = native.largest_tet_component(torch.tensor(tets, device=DEV), n)
    rows = np.concatenate([tets[:, 0], tets[:, 1], tets[:, 2], tets[:, 3]])
    cols = np.concatenate([tets[:, 1], tets[:, 2], tets[:, 3], tets[:, 0]])
    A = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
    nref, lref = csg.connected_components(A, directed=False)
    assert ncomp == nref
    lab = labels.cpu().numpy()
    for c in np.unique(lref):
        members = np.nonzero(lref == c)[0]
        assert np.all(lab[members] == members.min())
    sizes = np.bincount(lref)
    big = int(np.argmax(sizes))                 # first maximum = the component SciPy labelled first
    assert np.array_equal(kept.cpu().numpy(), np.nonzero(lref == big)[0])
    new = -np.ones(n, dtype=np.int64)
    new[lref == big] = np.arange(sizes[big])
    tref = new[tets]
    tref = tref[(tref >= 0).all(1)]
    assert np.array_equal(tout.cpu().numpy(), tref)


def test_thickness_tick_matches_reference():
    """One optimisation step of experiments/thickness_train.py:42-88: eigen loss through the marching-tets mesh and its
    gradient to the thickness logits."""
    g = golden("marching_tets")
    geo = _geometry(16, 1.5, g["tick_sdf"])
    with torch.no_grad():
        geo.marching_tets.thickness_coef.probablity.copy_(torch.tensor(g["tick_logits"]))
    assert abs(float(geo.get_thickness()) - float(g["tick_thickness"])) <= 1e-6
    target = geo.get_eigenvalues(thickness_coef=1.0)
    assert (np.abs(target.cpu().numpy() - g["tick_target"]) / g["tick_target"]).max() <= 2e-6
    loss = geo.tick(torch.tensor(g["tick_target"], device=DEV), 0, geo.FLAGS)
    assert abs(loss.item() - float(g["tick_loss"])) <= 1e-5 * float(g["tick_loss"])
    loss.backward()
    got = geo.marching_tets.thickness_coef.probablity.grad.numpy()
    ref = g["tick_grad_logits"]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-4
    assert len(list(geo.parameters())) == 1


def test_thickness_candidates_on_grid64_are_consistent():
    """BASELINE configs[3] geometry: 64-grid x 1.5, sphere SDF r = 0.6, shell 0 < sdf <= t max(sdf).  Structural checks at
    the size the sweep bench runs: one component after extraction, positive volumes summing to the shell's volume."""
    v, t = _grid(64)
    sdf = 0.6 - np.linalg.norm(v * 1.5, axis=1)
    geo = _geometry(64, 1.5, sdf.astype(np.float32), mode_num=32, order=2)
    for coef in (0.2, 0.9):
        _, _, av, at = geo.marching_tets(geo.verts, geo.sdf, geo.indices, coef)
        lv, lt = geo.get_largest_connected_component(av, at)
        assert geo.last_n_components == 1 and lt.shape == at.shape
        p = lv[lt].double()
        vol = torch.einsum("ij,ij->i", p[:, 1] - p[:, 0], torch.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0], dim=1)).abs() / 6
        th = coef * float(geo.marching_tets.max_thickness)
        shell = 4 / 3 * np.pi * (0.6 ** 3 - (0.6 - th) ** 3)
        assert abs(float(vol.sum()) - shell) <= 0.02 * shell


def test_sliver_mesh_eigenvalues_match_arpack():
    """Marching-tets output has sliver elements (sigma_min / sigma_max down to 3e-3 here): the solver must detect them, use
    the FP64 preconditioner and still meet the 1e-6 eigenvalue bar against SciPy ARPACK on the oracle's K, M."""
    from oracle import modal_oracle as mo
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    from diffsound_b200.diffelastic.material_model import MatSet
    g = golden("marching_tets")
    v, t = torch.tensor(g["g32_sphere_lcc_verts"]), torch.tensor(g["g32_sphere_lcc_tets"])
    rho, E, nu = MatSet.Steel[:3]
    K, M = mo.assemble(v, t, 1, E, nu, rho)
    lam, _, _, _ = mo.eig_arpack(K, M, 16)
    obj = DiffSoundObj(v.to(DEV), t.to(DEV), mode_num=16, order=1, mat=MatSet.Steel)
    assert obj._has_slivers()
    obj.eigen_decomposition()
    assert obj.eig_stats.get("precond") == "fp64"
    got = obj.eigenvalues.cpu().numpy()
    assert (np.abs(got - lam) / lam).max() <= 1e-6
