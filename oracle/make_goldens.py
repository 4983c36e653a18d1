"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the
UNMODIFIED reference (through oracle/ref_harness.py) on CPU in the build
container.  The GPU box has no /root/reference, so these files are the pins
for both the oracle (tests/test_oracle_vs_golden.py) and the CUDA path.

    python -m oracle.make_goldens            # writes tests/golden/

Cases:
  cube2, cube3   tiny Kuhn cubes: full K/M values, full U
  grid16         data/tets/16_tets.npz used directly as a tet mesh (config 1/3 style)
  bowl           data/mesh/bowl/bowl.obj_.msh (material experiments, configs 1-2)
each at order 1 and 2; plus oscillator cases.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from oracle import modal_oracle as mo         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NSAMPLE = 4096


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_meshes():
    meshes = {}
    for N in (2, 3):
        v, t = mo.kuhn_cube(N)
        meshes[f"cube{N}"] = (v.numpy(), t.numpy())
    d = np.load(rh.ref_path("data/tets/16_tets.npz"))
    meshes["grid16"] = (d["vertices"].astype(np.float32), d["indices"].astype(np.int64))
    pts, cells = rh.read_gmsh22(rh.ref_path("data/mesh/bowl/bowl.obj_.msh"))
    # mesh.py:52-53: torch.Tensor(points) (fp32), torch.Tensor(cells).long()
    meshes["bowl"] = (pts.astype(np.float32), cells["tetra"].astype(np.int64))
    return meshes


def modal_case(R, name, verts, tets, order, mat, k, full, with_grad=True):
    dm = R.diff_model
    torch.manual_seed(0)
    leaf = torch.tensor(verts).clone().requires_grad_(with_grad)
    obj = dm.DiffSoundObj(leaf, torch.tensor(tets).long(), mode_num=k, order=order, mat=mat)
    obj.eigen_decomposition()
    vals = obj.get_vals()
    g = (1.0 / obj.eigenvalues).float()        # upstream grad of a relative loss
    if with_grad:
        (vals[:, 0] * g).sum().backward()
    K = obj.stiff_matrix.detach()
    M = obj.mass_matrix.detach()
    n = K.shape[0]
    idx = K.indices().numpy()
    crow = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))]).astype(np.int64)
    col = idx[1].astype(np.int64)
    kv = K.values().numpy()
    mv = M.values().numpy()
    assert np.array_equal(M.indices().numpy(), idx)
    rng = np.random.default_rng(1234)
    samp = np.sort(rng.choice(kv.size, size=min(NSAMPLE, kv.size), replace=False))
    out = dict(
        order=order, k=k, material=np.array(mat, dtype=np.float64),
        n_nodes=obj.tetmesh.vertices.shape[0], nnz=kv.size,
        pverts_sha=sha(obj.tetmesh.vertices.detach().numpy()),
        ptets_sha=sha(obj.tetmesh.tets.numpy().astype(np.int64)),
        crow_sha=sha(crow), col_sha=sha(col),
        sample_idx=samp, K_sample=kv[samp], M_sample=mv[samp],
        K_absmax=np.abs(kv).max(), K_sum=kv.sum(), K_fro=np.sqrt((kv ** 2).sum()),
        M_absmax=np.abs(mv).max(), M_sum=mv.sum(), M_fro=np.sqrt((mv ** 2).sum()),
        eigenvalues=obj.eigenvalues.numpy(),
        get_vals=vals.detach().numpy(),
        upstream=g.numpy(),
        grad_verts=leaf.grad.numpy() if with_grad else np.zeros(0, np.float32),
    )
    if full or name in ("grid16",):
        out["pverts"] = obj.tetmesh.vertices.detach().numpy()
        out["ptets"] = obj.tetmesh.tets.numpy().astype(np.int32)
    if full:
        out.update(crow=crow, col=col, K_values=kv, M_values=mv)
    if full or (name == "grid16" and order == 1):
        out["U_hat"] = obj.U_hat.numpy()
    return out, obj


def material_case(R, name, verts, tets, order, mat, k):
    """task='material': trainable E, nu (diff_model.py:51-96, 371-388)."""
    dm = R.diff_model
    torch.manual_seed(7)
    obj = dm.DiffSoundObj(torch.tensor(verts), torch.tensor(tets).long(), mode_num=k, order=order,
                          mat=mat, mat_model=dm.TrainableLinear, task="material")
    mm = obj.material_model
    logits0 = (mm.youngs.probablity.detach().clone().numpy(), mm.poisson.probablity.detach().clone().numpy())
    E0, nu0 = float(mm.youngs()), float(mm.poisson())
    obj.eigen_decomposition()
    lam0 = obj.eigenvalues.numpy().copy()
    f0 = obj.get_undamped_freqs().detach().numpy()
    # a few "training steps" later the logits have moved but U, lambda are stale
    with torch.no_grad():
        mm.youngs.probablity += 0.1 * torch.randn(16)
        mm.poisson.probablity += 0.1 * torch.randn(16)
    logits1 = (mm.youngs.probablity.detach().clone().numpy(), mm.poisson.probablity.detach().clone().numpy())
    f1 = obj.get_undamped_freqs()
    w = torch.linspace(1.0, 2.0, k).unsqueeze(1)
    (f1 * w / f1.detach()).sum().backward()
    return dict(
        order=order, k=k, material=np.array(mat, dtype=np.float64),
        youngs_list=mm.youngs_list.numpy(), poisson_list=mm.poisson_list.numpy(),
        youngs_logits0=logits0[0], poisson_logits0=logits0[1], E0=E0, nu0=nu0,
        eigenvalues0=lam0, freqs0=f0,
        youngs_logits1=logits1[0], poisson_logits1=logits1[1],
        E1=float(mm.youngs()), nu1=float(mm.poisson()),
        freqs1=f1.detach().numpy(), weights=w.numpy(),
        grad_youngs_logits=mm.youngs.probablity.grad.numpy(),
        grad_poisson_logits=mm.poisson.probablity.grad.numpy(),
    )


def oscillator_cases(R, freqs):
    osc = R.oscillator
    Material = R.material_model.Material
    MatSet = R.material_model.MatSet
    out = {}
    # TraditionalDampedOscillator, config-2 shape: 1 audio, 16 modes, 8000 samples, 32 kHz, 150-tap impulse
    k, T, sr, F = 16, 8000, 32000, 150
    force = torch.zeros(1, F)
    force[0, 0] = 1
    f = torch.tensor(freqs[:k], dtype=torch.float32).reshape(k, 1)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        o = osc.TraditionalDampedOscillator(force.to(dt), 1, k, T, sr, Material(MatSet.Ceramic))
        fi = f.to(dt).clone().requires_grad_(True)
        y = o(fi)
        (0.5 * (y ** 2).sum()).backward()
        out[f"trad_audio_{tag}"] = y.detach().numpy()
        out[f"trad_gradf_{tag}"] = fi.grad.numpy()
        out[f"trad_damped_freq_{tag}"] = o.damped_freq[:, :, 0].detach().numpy()
    out["trad_freq"] = f.numpy()
    out["trad_meta"] = np.array([k, T, sr, F], dtype=np.int64)
    # a non-trivial force (decaying random taps) on 2 audios
    torch.manual_seed(3)
    force2 = (torch.randn(2, F) * torch.exp(-torch.arange(F) / 20.0)).double()
    o = osc.TraditionalDampedOscillator(force2, 2, k, T, sr, Material(MatSet.Glass))
    out["trad2_force"] = force2.numpy()
    out["trad2_audio_f64"] = o(f.double()).detach().numpy()
    # DampedOscillator.forward: learnable amp (B,k,1), alpha/beta (1,k,1) (oscillator.py:113-141)
    torch.manual_seed(5)
    B = 3
    forceB = torch.zeros(B, F, dtype=torch.float64)
    forceB[:, 0] = 1
    o = osc.DampedOscillator(forceB, B, k, T, sr, [0.0, 1.0], Material(MatSet.Ceramic)).double()
    o.alpha.values_list = o.alpha.values_list.double()
    o.beta.values_list = o.beta.values_list.double()
    fi = f.double().clone().requires_grad_(True)
    y = o(fi)
    (0.5 * (y ** 2).sum()).backward()
    out["damped_audio_f64"] = y.detach().numpy()
    out["damped_amp_param"] = o.amp.value.detach().numpy()
    out["damped_alpha_param"] = o.alpha.params.detach().numpy()
    out["damped_beta_param"] = o.beta.params.detach().numpy()
    out["damped_alpha_list"] = o.alpha.values_list.numpy()
    out["damped_beta_list"] = o.beta.values_list.numpy()
    out["damped_gradf"] = fi.grad.numpy()
    out["damped_grad_amp"] = o.amp.value.grad.numpy()
    out["damped_grad_alpha"] = o.alpha.params.grad.numpy()
    out["damped_grad_beta"] = o.beta.params.grad.numpy()
    out["damped_damped_freq"] = o.damped_freq[0, :, 0].detach().numpy()
    return out


def main():
    assert rh.reference_available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    R = rh.ref_modules()
    MatSet = R.material_model.MatSet
    meshes = load_meshes()
    np.savez_compressed(os.path.join(OUT, "meshes.npz"),
                        **{f"{n}_verts": v for n, (v, t) in meshes.items()},
                        **{f"{n}_tets": t.astype(np.int32) for n, (v, t) in meshes.items()})
    freqs_for_osc = None
    plan = [("cube2", 1, MatSet.Steel, 6, True), ("cube2", 2, MatSet.Steel, 8, True),
            ("cube3", 1, MatSet.Steel, 16, True), ("cube3", 2, MatSet.Steel, 16, True),
            ("grid16", 1, MatSet.Steel, 16, False), ("grid16", 2, MatSet.Steel, 16, False),
            ("bowl", 1, MatSet.Ceramic, 16, False), ("bowl", 2, MatSet.Ceramic, 16, False)]
    for name, order, mat, k, full in plan:
        v, t = meshes[name]
        print("modal", name, order, flush=True)
        dst = os.path.join(OUT, f"modal_{name}_o{order}.npz")
        if os.path.exists(dst) and "--force" not in sys.argv:
            if name == "bowl" and order == 2:
                freqs_for_osc = np.sqrt(np.load(dst)["eigenvalues"]) / 2 / np.pi
            continue
        # bowl order 2 with autograd needs > 62 GB in the reference (T*G*(3N)^2 COO triples
        # kept alive per batch); its gradient golden is skipped.
        out, obj = modal_case(R, name, v, t, order, mat, k, full, with_grad=not (name == "bowl" and order == 2))
        np.savez_compressed(dst, **out)
        if name == "bowl" and order == 2:
            freqs_for_osc = np.sqrt(out["eigenvalues"]) / 2 / np.pi
    for name, order in (("cube3", 2), ("grid16", 1), ("bowl", 1)):
        v, t = meshes[name]
        print("material", name, order, flush=True)
        if os.path.exists(os.path.join(OUT, f"material_{name}_o{order}.npz")) and "--force" not in sys.argv:
            continue
        out = material_case(R, name, v, t, order, MatSet.Ceramic, 16)
        np.savez_compressed(os.path.join(OUT, f"material_{name}_o{order}.npz"), **out)
    print("oscillator", flush=True)
    np.savez_compressed(os.path.join(OUT, "oscillator.npz"), **oscillator_cases(R, freqs_for_osc))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
