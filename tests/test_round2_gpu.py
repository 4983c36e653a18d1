"""GPU parity tests added in round 2: rows the round-1 verdict listed as untested or missing.

Goldens come from the unmodified reference (oracle/make_goldens_r2.py).  Tolerances:
  mat_baseline freqs <= 2e-6 rel; d f / d(youngs logits) <= 2e-4 rel-L2 -- the reference evaluates
      U^T K(theta) U - lambda U^T M U in fp32 on values ~1e9 (diff_model.py:382-386), so its own gradient carries
      ~1e-4 of rounding noise; our value is the fp64 one
  stiff_func values / gradients <= 2e-5 rel-L2 (the reference path is fp32 end to end, deform.py:70-165)
  oscillator variants / filtered noise / audio <= 1e-4 rel-L2 (north star)
  spectral losses <= 1e-4 relative on the value, <= 1e-2 rel-L2 on the gradient (fp32 FFTs on both sides; the L1
      terms contribute sign(a - b) per bin, which flips wherever the two fp32 spectra differ by rounding, and 1 / (S + eps)
      amplifies fp32 noise in quiet bins: measured 3.3e-3 (l1) and 3.1e-3 (rmse); the kernel's own gradient is checked
      against central differences of its fp64-accumulated forward in test_mss_gradient_finite_difference)
  64 modes: eigenvalues <= 1e-6 relative
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


# ------------------------------------------------------------------------------------------------
# task = "mat_baseline"  (BASELINE configs[0]: fixed Poisson ratio)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,order", [("grid16", 1), ("bowl", 1), ("cube3", 2)])
def test_mat_baseline_matches_reference(meshes, name, order):
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj, TrainableLinear
    g = golden(f"baseline_{name}_o{order}")
    v, t = meshes[name]
    obj = DiffSoundObj(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV).long(), mode_num=int(g["k"]), order=order,
                       mat=tuple(g["material"]), mat_model=TrainableLinear, task="mat_baseline")
    mm = obj.material_model
    assert mm.poisson_list.shape == (1,) and np.allclose(mm.poisson_list.numpy(), g["poisson_list"])
    params = list(obj.parameters())
    assert len(params) == 1 and params[0] is mm.youngs.probablity           # diff_model.py:149-150
    with torch.no_grad():
        mm.youngs.probablity.copy_(torch.tensor(g["youngs_logits1"]))
        mm.poisson.probablity.copy_(torch.tensor(g["poisson_logits"]))
    # the reference decomposed at logits0 and evaluated at logits1; lambda(E) is linear in E at fixed nu, so
    # decomposing at logits1 gives the same frequencies (the first-order term is exact)
    obj.eigen_decomposition()
    f1 = obj.get_undamped_freqs()
    assert (np.abs(f1.detach().cpu().numpy() - g["freqs1"]) / g["freqs1"]).max() <= 2e-6
    w = torch.tensor(g["weights"], device=DEV)
    (f1 * w / f1.detach()).sum().backward()
    assert rel(mm.youngs.probablity.grad.numpy(), g["grad_youngs_logits"]) <= 2e-4


# ------------------------------------------------------------------------------------------------
# stiff_func values and gradients
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,order", [("cube3", 2), ("grid16", 1)])
def test_stiff_func_values_match_reference(meshes, name, order):
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj, TrainableLinear
    g = golden(f"stiff_{name}_o{order}")
    v, t = meshes[name]
    obj = DiffSoundObj(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV).long(), mode_num=16, order=order,
                       mat=tuple(g["material"]), mat_model=TrainableLinear, task="material")
    mm = obj.material_model
    with torch.no_grad():
        mm.youngs.probablity.copy_(torch.tensor(g["youngs_logits"]))
        mm.poisson.probablity.copy_(torch.tensor(g["poisson_logits"]))
    X = torch.tensor(g["X"], device=DEV).requires_grad_(True)
    Y = obj.stiff_func(X)
    assert Y.shape == X.shape and Y.dtype == torch.float32
    assert rel(Y.detach().cpu().numpy(), g["Y"]) <= 2e-5
    (Y * torch.tensor(g["G"], device=DEV)).sum().backward()
    assert rel(X.grad.cpu().numpy(), g["grad_X"]) <= 2e-5
    assert rel(mm.youngs.probablity.grad.numpy(), g["grad_youngs_logits"]) <= 1e-4
    assert rel(mm.poisson.probablity.grad.numpy(), g["grad_poisson_logits"]) <= 1e-4
    y1 = obj.stiff_func(torch.tensor(g["x1"], device=DEV))
    assert y1.shape == (X.shape[0],) and rel(y1.detach().cpu().numpy(), g["y1"]) <= 2e-5
    # more than 128 columns (the reference has no limit)
    Xw = torch.randn(X.shape[0], 130, device=DEV)
    Yw = obj.stiff_func(Xw)
    assert Yw.shape == Xw.shape
    assert torch.allclose(Yw[:, 129], obj.stiff_func(Xw[:, 129]), rtol=1e-5, atol=1e-5 * float(Yw.abs().max()))


# ------------------------------------------------------------------------------------------------
# oscillator variants + filtered noise
# ------------------------------------------------------------------------------------------------
def test_forward_curve_and_early_match_reference():
    from scipy import interpolate
    from diffsound_b200.ddsp import oscillator as osc
    from diffsound_b200.diffelastic.material_model import Material, MatSet
    g = golden("oscillator_r2")
    k, T, sr, F, B = (int(x) for x in g["meta"])
    f = torch.tensor(g["freq"], device=DEV)
    force = torch.tensor(g["force"], dtype=torch.float32, device=DEV)
    o = osc.DampedOscillator(force, B, k, T, sr, [20, 16000], Material(MatSet.Ceramic))
    curve = interpolate.interp1d(g["curve_x"], g["curve_y"], fill_value="extrapolate")
    fi = f.clone().requires_grad_(True)
    y = o.forward_curve(fi, curve)
    assert y.shape == (B, T) and y.dtype == torch.float32
    assert rel(y.detach().cpu().numpy(), g["curve_audio_f64"]) <= 1e-4
    assert o.damped_freq.shape == tuple(g["curve_damped_freq"].shape)
    assert np.allclose(o.damped_freq.detach().cpu().numpy(), g["curve_damped_freq"], rtol=1e-6)
    (0.5 * (y ** 2).sum()).backward()
    assert rel(fi.grad.cpu().numpy(), g["curve_gradf"]) <= 1e-3       # fp32 phase at 8000 samples
    e = o.early(f, curve)
    assert rel(e.detach().cpu().numpy(), g["early_audio_f64"]) <= 1e-4
    # state-dict parity with the reference module (names and shapes)
    sd = o.state_dict()
    assert sorted(sd.keys()) == [str(s) for s in g["damped_state_keys"]]
    assert [str(tuple(sd[n].shape)) for n in sorted(sd.keys())] == [str(s) for s in g["damped_state_shapes"]]
    # noise_rate / non_linear_rate are accepted and ignored by DampedOscillator.forward (oscillator.py:113-141)
    a = o(f, non_linear_rate=0.3, noise_rate=2e-4)
    assert torch.equal(a, o(f))


def _load_gt(g, prefix="gt_"):
    from diffsound_b200.ddsp import oscillator as osc
    from diffsound_b200.diffelastic.material_model import Material, MatSet
    k, T, sr, F, B = (int(x) for x in g[prefix + "meta"]) if prefix + "meta" in g.files else (None,) * 5
    return osc, Material, MatSet, k, T, sr, F, B


def test_gt_oscillator_with_noise_matches_reference():
    g = golden("oscillator_r2")
    osc, Material, MatSet, k, T, sr, F, B = _load_gt(g)
    force = torch.zeros(B, F, device=DEV)
    force[:, 0] = 1
    o = osc.GTDampedOscillator(force, B, k, T, sr, [20, 16000], Material(MatSet.Ceramic)).cuda()
    sd = o.state_dict()
    assert sorted(sd.keys()) == [str(s) for s in g["gt_state_keys"]]
    assert [str(tuple(sd[n].shape)) for n in sorted(sd.keys())] == [str(s) for s in g["gt_state_shapes"]]
    with torch.no_grad():
        o.freq_linear.params.copy_(torch.tensor(g["gt_freq_params"], dtype=torch.float32))
        o.alpha.params.copy_(torch.tensor(g["gt_alpha_params"], dtype=torch.float32))
        o.beta.params.copy_(torch.tensor(g["gt_beta_params"], dtype=torch.float32))
        o.amp.value.copy_(torch.tensor(g["gt_amp_value"], dtype=torch.float32))
        o.noise.coefficient_bank.copy_(torch.tensor(g["gt_noise_bank"], dtype=torch.float32))
    y0 = o(noise_rate=0.0)
    assert y0.shape == (B, T)
    assert rel(y0.detach().cpu().numpy(), g["gt_audio_nonoise_f64"]) <= 1e-4
    assert np.allclose(o.undamped_freq.detach().cpu().numpy(), g["gt_undamped_freq"], rtol=2e-6)
    assert np.allclose(o.damping().detach().cpu().numpy(), g["gt_damping"], rtol=2e-6)
    torch.manual_seed(int(g["gt_noise_seed"]))
    y1 = o(noise_rate=2e-4)
    assert rel(y1.detach().cpu().numpy(), g["gt_audio_noise_f64"]) <= 1e-4
    # the noise really is in there
    assert rel(y1.detach().cpu().numpy(), g["gt_audio_nonoise_f64"]) > 1e-6
    (0.5 * (y1 ** 2).sum()).backward()
    for got, key, tol in ((o.freq_linear.params.grad, "gt_grad_freq_params", 2e-3), (o.alpha.params.grad, "gt_grad_alpha_params", 1e-3),
                          (o.beta.params.grad, "gt_grad_beta_params", 1e-3), (o.amp.value.grad, "gt_grad_amp_value", 1e-4),
                          (o.noise.coefficient_bank.grad, "gt_grad_noise_bank", 1e-3)):
        assert rel(got.cpu().numpy(), g[key]) <= tol, key
    with pytest.raises(NotImplementedError):
        o(non_linear_rate=0.1)


def test_filtered_noise_matches_reference():
    from diffsound_b200.ddsp.filtered_noise import FilteredNoise
    g = golden("oscillator_r2")
    fn = FilteredNoise(2, 1000).cuda()
    assert tuple(fn.coefficient_bank.shape) == tuple(g["fn_bank"].shape)
    with torch.no_grad():
        fn.coefficient_bank.copy_(torch.tensor(g["fn_bank"]))
    torch.manual_seed(int(g["fn_seed"]))
    y = fn()
    assert y.shape == (2, 1000) and y.dtype == torch.float32
    assert rel(y.detach().cpu().numpy(), g["fn_audio"]) <= 1e-5
    (y * torch.tensor(g["fn_weight"], device=DEV)).sum().backward()
    assert rel(fn.coefficient_bank.grad.cpu().numpy(), g["fn_grad_bank"]) <= 1e-4


# ------------------------------------------------------------------------------------------------
# multi-scale spectral loss
# ------------------------------------------------------------------------------------------------
def test_mss_loss_matches_reference():
    from diffsound_b200.ddsp.mss_loss import MSSLoss, SSSLoss
    g = golden("mss_loss")
    pred = torch.tensor(g["pred"], device=DEV)
    true = torch.tensor(g["true"], device=DEV)
    sr = int(g["sample_rate"])
    for tag, typ in (("l1_a", "l1_loss"), ("l1_b", "l1_loss"), ("rmse", "rmse_loss"), ("l1_2048", "l1_loss")):
        ffts = [int(x) for x in g[f"{tag}_nffts"]]
        lf = MSSLoss(ffts, sr, type=typ).cuda()
        per = [float(l(pred, true)) for l in lf.losses]
        assert np.allclose(per, g[f"{tag}_per_scale"], rtol=1e-4), (tag, per, g[f"{tag}_per_scale"])
        p = pred.clone().requires_grad_(True)
        loss = lf(p, true)
        assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 1e-4 * abs(float(g[f"{tag}_loss"])), tag
        loss.backward()
        assert rel(p.grad.cpu().numpy(), g[f"{tag}_grad"]) <= 1e-2, tag
    s = SSSLoss(256, sr, type="l1_loss")
    S = s.spec(pred).cpu().numpy()
    assert S.shape == g["spec256"].shape
    assert np.abs(S - g["spec256"]).max() <= 1e-5 * np.abs(g["spec256"]).max()
    ls = s.log_spec(pred[0]).cpu().numpy()
    assert ls.shape == g["logspec256"].shape and np.abs(ls - g["logspec256"]).max() <= 2e-2   # log2 near eps amplifies fp32 noise
    with pytest.raises(NotImplementedError):
        MSSLoss([1024], sr, type="geomloss")(pred, true)


def test_mss_gradient_finite_difference():
    """gradient of both loss types against central differences of the kernel's own forward (fp64 loss value).  Both
    signals carry a strong noise floor: log2(S + eps) is only linear over a finite step where S is far above both eps
    and the step's own spectrum (checked on the CPU with torch.stft autograd: step 2e-4 agrees to 3e-4)."""
    from diffsound_b200 import native
    torch.manual_seed(0)
    T, n_fft, hop = 700, 128, 32
    t = torch.arange(T, device=DEV) / 8000.0
    true = (torch.sin(2 * np.pi * 440 * t) * torch.exp(-6 * t)).reshape(1, T).float() + 0.5 * torch.randn(1, T, device=DEV)
    pred = (0.8 * torch.sin(2 * np.pi * 470 * t + 0.3) * torch.exp(-5 * t)).reshape(1, T).float() + 0.5 * torch.randn(1, T, device=DEV)
    d = torch.randn(1, T, device=DEV)
    for mode in (1, 0):
        loss, scratch = native.mss_loss_fwd(pred, true, n_fft, hop, mode, 1.0, 1e-7)
        gx = torch.empty_like(pred)
        native.mss_loss_bwd(pred, true, n_fft, hop, mode, 1.0, 1e-7, loss, 1.0, scratch, gx, False)
        an = float((gx.double() * d.double()).sum())
        h = 2e-4
        lp, _ = native.mss_loss_fwd((pred + h * d).contiguous(), true, n_fft, hop, mode, 1.0, 1e-7)
        lm, _ = native.mss_loss_fwd((pred - h * d).contiguous(), true, n_fft, hop, mode, 1.0, 1e-7)
        fd = (float(lp) - float(lm)) / (2 * h)
        assert abs(fd - an) <= 3e-2 * abs(fd) + 1e-6, (mode, fd, an)


# ------------------------------------------------------------------------------------------------
# more modes than one eigensolver block holds (geometry_train.py:147: mode_num = 64)
# ------------------------------------------------------------------------------------------------
def test_mode_num_64_matches_arpack(meshes):
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    g = golden("modal_bowl_o1_k64")
    v, t = meshes["bowl"]
    leaf = torch.tensor(v, device=DEV).requires_grad_(True)
    obj = DiffSoundObj(leaf, torch.tensor(t, device=DEV).long(), mode_num=64, order=1, mat=tuple(g["material"]))
    obj.eigen_decomposition()
    lam = obj.eigenvalues.cpu().numpy()
    assert lam.shape == (64,) and np.all(np.diff(lam) >= 0)
    assert (np.abs(lam - g["eigenvalues"]) / g["eigenvalues"]).max() <= 1e-6
    U = obj.U_hat
    assert U.shape == (obj.deform.pattern.n, 64) and obj.U_hat_full.shape[1] == 70
    M = obj.mass_matrix
    assert float((U.T @ (M @ U) - torch.eye(64, device=DEV, dtype=torch.float64)).abs().max()) <= 1e-8
    assert obj.eig_stats["batches"] == 2
    vals = obj.get_vals()
    assert vals.shape == (64, 1)
    assert (np.abs(vals.detach().cpu().numpy() - g["get_vals"]) / g["get_vals"]).max() <= 1.2e-6
    (vals[:, 0] * torch.tensor(g["upstream"], device=DEV)).sum().backward()
    # bowl pairs (split by ~1 %) are resolved, the sum over all 64 modes is basis independent unless the 64th
    # mode sits in a cluster with the 65th; the golden's own spectrum says it does not when this holds
    assert rel(leaf.grad.cpu().numpy(), g["grad_verts"]) <= 1e-5
    f = obj.get_undamped_freqs()
    assert f.shape == (64, 1)


def test_backward_uses_the_eigenpairs_of_its_forward(meshes):
    """get_vals() -> another eigen_decomposition() (warm start overwrote the block in place) -> backward
    must give the gradient of the FIRST decomposition's eigenpairs."""
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    g = golden("modal_grid16_o1")
    v, t = meshes["grid16"]
    up = torch.tensor(g["upstream"], device=DEV)

    def run(perturb):
        leaf = torch.tensor(v, device=DEV).requires_grad_(True)
        obj = DiffSoundObj(leaf, torch.tensor(t, device=DEV).long(), mode_num=16, order=1, mat=tuple(g["material"]))
        obj.eigen_decomposition()
        vals = obj.get_vals()
        u_before = obj.U_hat.clone()
        if perturb:
            obj.material_model.youngs = obj.material_model.youngs * 1.7      # different K: different eigenpairs
            obj.eigen_decomposition()                                        # warm start from the old block
            assert torch.equal(u_before, ctx_u(vals))                        # the pinned block was not overwritten
        (vals[:, 0] * up).sum().backward()
        return leaf.grad.clone(), u_before

    def ctx_u(vals):
        return vals.grad_fn.U

    g0, _ = run(False)
    g1, _ = run(True)
    assert torch.equal(g0, g1)


# ------------------------------------------------------------------------------------------------
# the reference's experiment loops through the `src.*` alias
# ------------------------------------------------------------------------------------------------
@pytest.fixture()
def src_alias():
    import diffsound_b200
    names = ["", ".diffelastic", ".diffelastic.diff_model", ".diffelastic.mesh", ".ddsp", ".ddsp.oscillator", ".ddsp.mss_loss",
             ".ddsp.filtered_noise", ".lobpcg", ".cuda_module"]
    saved = {}
    for n in names:
        mod = importlib.import_module("diffsound_b200" + n)
        saved["src" + n] = sys.modules.get("src" + n)
        sys.modules["src" + n] = mod
    yield
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def _write_bowl(tmp_path, meshes):
    from diffsound_b200.diffelastic.mesh import write_msh
    v, t = meshes["bowl"]
    base = os.path.join(str(tmp_path), "bowl.obj")
    write_msh(base + "_.msh", v.astype(np.float64), t, "tetra")
    return base


def test_material_sync_train_inner_step(src_alias, meshes, tmp_path):
    """experiments/material_sync_train.py:95-168 (one material, exp_mode 2, late loss) with the script's own imports
    resolved through the alias; loss, RMSE, gradients and the Adam update against the reference's values."""
    from src.diffelastic.diff_model import Material, build_model, MatSet          # noqa: F401  (script line 14)
    from src.ddsp.mss_loss import MSSLoss                                         # script line 19
    from src.ddsp.oscillator import TraditionalDampedOscillator                   # script line 20
    from torch.optim import Adam, lr_scheduler
    g = golden("step_material_sync")
    sample_rate, frame_num, force_frame_num, eigen_num = (int(x) for x in g["meta"])
    mesh_dir = _write_bowl(tmp_path, meshes)
    gt_material_coeff = [float(x) for x in g["gt_coeff"]]
    material_coeff = [float(x) for x in g["init_coeff"]]
    gt_forces = torch.zeros((1, force_frame_num)).cuda()
    gt_forces[0, 0] = 1
    gt_osc = TraditionalDampedOscillator(gt_forces, 1, eigen_num, frame_num, sample_rate, Material(gt_material_coeff)).cuda()
    gt_model = build_model(mesh_dir, mode_num=eigen_num, order=2, mat=gt_material_coeff, task="gt")
    gt_model.eigen_decomposition()
    gt_undamped_freq = gt_model.get_undamped_freqs().float()
    assert (np.abs(gt_undamped_freq.cpu().numpy() - g["gt_freq"]) / g["gt_freq"]).max() <= 2e-6
    gt_audios = gt_osc(gt_undamped_freq)
    assert rel(gt_audios.cpu().numpy(), g["gt_audios"]) <= 1e-4
    model = build_model(mesh_dir, mode_num=eigen_num, order=1, mat=material_coeff, task="material")
    with torch.no_grad():                       # the reference's logits after its own (seeded) pre-training
        model.material_model.youngs.probablity.copy_(torch.tensor(g["youngs_logits0"]))
        model.material_model.poisson.probablity.copy_(torch.tensor(g["poisson_logits0"]))
    oscillator = TraditionalDampedOscillator(gt_forces, len(gt_audios), eigen_num, frame_num, sample_rate,
                                             Material(material_coeff)).cuda()
    late_loss_func = MSSLoss([1024, 512, 256, 128, 64], sample_rate, type='l1_loss').cuda()
    rmse_loss_func = MSSLoss([1024, 512, 256, 128, 64], sample_rate, type='rmse_loss').cuda()
    optimizer_model = Adam(model.parameters(), lr=2e-3)
    scheduler_model = lr_scheduler.StepLR(optimizer_model, step_size=100, gamma=0.95)
    # ---- the loop body, lines 137-168
    model.eigen_decomposition()
    undamped_freq = model.get_undamped_freqs().float()
    assert (np.abs(undamped_freq.detach().cpu().numpy() - g["undamped_freq"]) / g["undamped_freq"]).max() <= 2e-6
    predict_signal = oscillator(undamped_freq)
    damped_freq = oscillator.damped_freq
    spec_scale = 1
    loss = late_loss_func(predict_signal, torch.tensor(g["gt_audios"], device=DEV), damped_freq, spec_scale)
    optimizer_model.zero_grad()
    loss.backward()
    mm = model.material_model
    gE, gnu = mm.youngs.probablity.grad.numpy().copy(), mm.poisson.probablity.grad.numpy().copy()
    optimizer_model.step()
    scheduler_model.step()
    with torch.no_grad():
        RMSE_loss = rmse_loss_func(predict_signal, torch.tensor(g["gt_audios"], device=DEV))
    # the reference renders its audio through fp32 cumsum phases (3-4e-5 rel-L2 from the closed form); the
    # log-spectral L1 terms amplify that in quiet bins: measured 4.2e-4 on the loss
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    # RMSE of log2(S + 1e-7): dominated by the quiet bins, where the reference's fp32 cumsum noise IS the spectrum
    assert abs(RMSE_loss.item() - float(g["rmse"])) <= 5e-3 * float(g["rmse"])
    # Gradient of an L1 log-spectral loss of fp32 audio: the sign() terms of quiet bins flip with the rounding of the rendering.
    # The UNMODIFIED reference does not reproduce its own gradient better than ~1e-2 (oracle/sync_grad_sensitivity.py, CPU,
    # same inputs): fp32 run vs a second fp32 run 1.3e-3 (youngs logits) / 8.8e-3 (poisson logits); fp32 vs the same code in
    # float64 2.3e-3 .. 1.1e-2 / 4.1e-3 .. 1.7e-2; one fp32 run vs the committed golden 1.5e-2 -- while the loss value agrees to
    # 4e-4 every time.  (E, nu) enter through one scalar each, so the logit gradients differ by a uniform factor.  Measured here:
    # 1.2e-2 / 3.3e-2 against the fp32 golden.  The differentiation itself is pinned elsewhere: d freq / d(E, nu) to 1e-5
    # (test_material_gradients_match_reference), the oscillator backward to 1e-3 and the MSS backward by finite differences.
    assert rel(gE, g["grad_youngs_logits"]) <= 5e-2 and rel(gnu, g["grad_poisson_logits"]) <= 5e-2
    g64 = golden("step_material_sync_f64")        # the same step of the reference evaluated in float64 (audio + loss)
    assert rel(gE, g64["grad_youngs_logits"]) <= 5e-2 and rel(gnu, g64["grad_poisson_logits"]) <= 5e-2
    assert abs(loss.item() - float(g64["loss"])) <= 1e-3 * float(g64["loss"])
    # first Adam step: every logit moves by lr * sign(grad) (bias-corrected m / sqrt(v) = +-1)
    assert np.allclose(mm.youngs.probablity.detach().numpy(), g["youngs_logits1"], atol=2e-5)
    assert np.allclose(mm.poisson.probablity.detach().numpy(), g["poisson_logits1"], atol=2e-5)


def test_material_real_train_inner_steps(src_alias, meshes, tmp_path):
    """experiments/material_real_train.py:113-132 (pre-oscillator step: GTDampedOscillator with noise_rate = 2e-4) and
    :176-205 (main-loop step through DampedOscillator.forward_curve)."""
    from scipy import interpolate
    from src.diffelastic.diff_model import Material, build_model, MatSet
    from src.ddsp.mss_loss import MSSLoss
    from src.ddsp.oscillator import DampedOscillator, GTDampedOscillator, init_damps   # noqa: F401
    from torch.optim import Adam
    g = golden("step_material_real")
    sample_rate, frame_num, force_frame_num, eigen_num, audio_num = (int(x) for x in g["meta"])
    material_coeff = getattr(MatSet, "Ceramic")
    gt_audios = torch.tensor(g["gt_audios"], device=DEV)
    gt_forces = torch.zeros((1, force_frame_num)).cuda()
    gt_forces[0, 0] = 1
    gt_forces = gt_forces.repeat(len(gt_audios), 1)
    late_loss_func = MSSLoss([512, 256, 128, 64, 32], sample_rate, type='l1_loss').cuda()
    pre_osc = GTDampedOscillator(gt_forces, len(gt_audios), eigen_num * 16, frame_num, sample_rate, [20, 16000],
                                 Material(material_coeff)).cuda()
    with torch.no_grad():
        pre_osc.freq_linear.params.copy_(torch.tensor(g["pre_freq_params"]))
        pre_osc.alpha.params.copy_(torch.tensor(g["pre_alpha_params"]))
        pre_osc.beta.params.copy_(torch.tensor(g["pre_beta_params"]))
        pre_osc.amp.value.copy_(torch.tensor(g["pre_amp_value"]))
        pre_osc.noise.coefficient_bank.copy_(torch.tensor(g["pre_noise_bank"]))
    optimizer_pre_osc = Adam(pre_osc.parameters(), lr=5e-3)
    torch.manual_seed(int(g["pre_noise_seed"]))
    predict_signal = pre_osc(noise_rate=2e-4)
    assert rel(predict_signal.detach().cpu().numpy(), g["pre_signal"]) <= 1e-4
    loss = late_loss_func(predict_signal, gt_audios)
    optimizer_pre_osc.zero_grad()
    loss.backward()
    assert abs(loss.item() - float(g["pre_loss"])) <= 2e-4 * float(g["pre_loss"])
    assert rel(pre_osc.amp.value.grad.cpu().numpy(), g["pre_grad_amp"]) <= 5e-3
    assert rel(pre_osc.freq_linear.params.grad.cpu().numpy(), g["pre_grad_freq"]) <= 2e-2
    assert rel(pre_osc.noise.coefficient_bank.grad.cpu().numpy(), g["pre_grad_noise"]) <= 5e-3
    optimizer_pre_osc.step()
    # ---- main loop step
    damping_curve = interpolate.interp1d(g["curve_x"], g["curve_y"], fill_value="extrapolate")
    mesh_dir = _write_bowl(tmp_path, meshes)
    model = build_model(mesh_dir, mode_num=eigen_num, order=1, mat=material_coeff, task="material")
    with torch.no_grad():
        model.material_model.youngs.probablity.copy_(torch.tensor(g["youngs_logits0"]))
        model.material_model.poisson.probablity.copy_(torch.tensor(g["poisson_logits0"]))
    oscillator = DampedOscillator(gt_forces, len(gt_audios), eigen_num, frame_num, sample_rate, f_range=[20, 16000],
                                  mat=Material(material_coeff)).cuda()
    late = MSSLoss([1024, 512, 256, 128, 64], sample_rate, type='l1_loss').cuda()
    optimizer_model = Adam(model.parameters(), lr=1e-3)
    model.eigen_decomposition()
    undamped_freq = model.get_undamped_freqs().float()
    assert (np.abs(undamped_freq.detach().cpu().numpy() - g["main_undamped"]) / g["main_undamped"]).max() <= 2e-6
    predict_signal = oscillator.forward_curve(undamped_freq, damping_curve)
    assert rel(predict_signal.detach().cpu().numpy(), g["main_predict"]) <= 1e-4
    loss = late(predict_signal, gt_audios, oscillator.damped_freq, 1)
    optimizer_model.zero_grad()
    loss.backward()
    assert abs(loss.item() - float(g["main_loss"])) <= 2e-4 * float(g["main_loss"])
    mm = model.material_model
    assert rel(mm.youngs.probablity.grad.numpy(), g["main_grad_youngs"]) <= 2e-2
    assert rel(mm.poisson.probablity.grad.numpy(), g["main_grad_poisson"]) <= 2e-2
    optimizer_model.step()


# ------------------------------------------------------------------------------------------------
# parity closer to the size the bench times
# ------------------------------------------------------------------------------------------------
def test_grid32_order2_matches_arpack():
    """data/tets/32_tets.npz at order 2 (n = 137 880, nnz = 10 948 230): pattern, sampled K / M values and the 32
    lowest elastic eigenvalues against the reference's ARPACK run."""
    from diffsound_b200.diffelastic.diff_model import DiffSoundObj
    import hashlib
    path = os.path.join(os.path.dirname(__file__), "golden", "modal_grid32_o2.npz")
    mesh = os.path.join(os.path.dirname(__file__), "golden", "mesh_grid32.npz")
    if not (os.path.exists(path) and os.path.exists(mesh)):
        pytest.skip("grid32 golden not generated")
    g, m = np.load(path), np.load(mesh)
    obj = DiffSoundObj(torch.tensor(m["verts"], device=DEV), torch.tensor(m["tets"].astype(np.int64), device=DEV),
                       mode_num=32, order=2, mat=tuple(g["material"]))
    obj.eigen_decomposition()
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    K = obj.stiff_matrix
    idx = K.indices().cpu().numpy()
    n = K.shape[0]
    assert n == 137880 and idx.shape[1] == int(g["nnz"])
    crow = np.concatenate([[0], np.cumsum(np.bincount(idx[0], minlength=n))]).astype(np.int64)
    assert sha(crow) == str(g["crow_sha"]) and sha(idx[1].astype(np.int64)) == str(g["col_sha"])
    kv, mv = K.values().cpu().numpy(), obj.mass_matrix.values().cpu().numpy()
    s = g["sample_idx"]
    assert np.abs(kv[s] - g["K_sample"]).max() <= 2e-6 * float(g["K_absmax"])
    assert np.allclose(mv[s], g["M_sample"], rtol=1e-12, atol=0)
    lam = obj.eigenvalues.cpu().numpy()
    assert (np.abs(lam - g["eigenvalues"]) / g["eigenvalues"]).max() <= 1e-6
    vals = obj.get_vals().cpu().numpy()
    assert (np.abs(vals - g["get_vals"]) / g["get_vals"]).max() <= 1.2e-6


# ------------------------------------------------------------------------------------------------
# lobpcg API: operator-form driver (callable A, iK, largest=True, tracker)
# ------------------------------------------------------------------------------------------------
def _dense_pencil(n=400, seed=5):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * np.geomspace(1.0, 1e3, n)) @ Q.T
    Lb = np.tril(rng.standard_normal((n, n)) * 0.05) + np.eye(n)
    B = Lb @ Lb.T
    A, B = (A + A.T) / 2, (B + B.T) / 2
    import scipy.linalg as sla
    w = sla.eigh(A, B, eigvals_only=True)
    return A, B, w


def test_lobpcg_callable_operator_and_preconditioner():
    """lobpcg_func(A callable, B, k, iK=..., largest=False) as _lobpcg.py:123-212 allows; tracker / force_stop hooks."""
    from diffsound_b200.lobpcg import lobpcg, lobpcg_func
    A, B, w = _dense_pencil()
    At, Bt = torch.tensor(A, device=DEV), torch.tensor(B, device=DEV)
    iK = torch.linalg.inv(At)
    k = 6
    calls = []

    def tracker(worker):
        calls.append((worker.ivars["istep"], worker.ivars.get("converged_count", 0)))

    torch.manual_seed(0)
    E, X, rerr = lobpcg_func(lambda V: At @ V, Bt, k, X=torch.randn(400, 8, device=DEV, dtype=torch.float64), iK=iK, niter=200,
                             tol=1e-10, largest=False, tracker=tracker, return_rerr=True)
    assert E.shape == (k,) and X.shape == (400, k) and E.dtype == torch.float64
    assert np.abs(E.cpu().numpy() - w[:k]).max() <= 1e-8 * w[k]
    G = X.T @ Bt @ X
    assert float((G - torch.eye(k, device=DEV, dtype=torch.float64)).abs().max()) <= 1e-8
    assert float((At @ X - Bt @ X * E).norm()) <= 1e-6 * float(E.abs().max())
    assert calls and calls[0][0] == 0 and calls[-1][1] >= k and rerr.shape == (8,)
    # preconditioner given as a callable, operators as fp32 tensors (the reference's own precision)
    E32, X32 = lobpcg(At.float(), k, Bt.float(), iK=lambda R: iK.to(R.dtype) @ R, niter=300, tol=1e-5, largest=False)
    assert E32.dtype == torch.float32 and np.abs(E32.cpu().numpy() - w[:k]).max() <= 1e-4 * w[k]
    # force_stop from the tracker (_lobpcg.py:336-342)
    steps = []

    def stopper(worker):
        steps.append(worker.ivars["istep"])
        if worker.ivars["istep"] >= 3:
            worker.bvars["force_stop"] = True

    lobpcg(At, k, Bt, niter=100, tol=1e-14, largest=False, tracker=stopper)
    assert max(steps) == 3


def test_lobpcg_largest_default():
    """`largest` defaults to True in the reference's signature (_lobpcg.py:64,179)."""
    from diffsound_b200.lobpcg import lobpcg
    A, B, w = _dense_pencil(300, seed=9)
    At, Bt = torch.tensor(A, device=DEV), torch.tensor(B, device=DEV)
    torch.manual_seed(1)
    E, X = lobpcg(At, 4, Bt, niter=2000, tol=1e-9)
    ref = w[::-1][:4]
    assert np.abs(E.cpu().numpy() - ref).max() <= 1e-6 * ref[0]
    assert float((At @ X - Bt @ X * E).norm()) <= 1e-4 * ref[0]
    with pytest.raises(ValueError):
        lobpcg(At[:20, :20], 10, Bt[:20, :20])          # m < 3 n (_lobpcg.py:42-46)
