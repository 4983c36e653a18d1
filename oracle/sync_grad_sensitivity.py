"""TEST INFRASTRUCTURE ONLY.  How far does the REFERENCE's own gradient of the material_sync_train inner step move when its
audio is rendered in float64 instead of float32 (fp32 cumsum phases, oscillator.py:297-298)?  The answer bounds what any
implementation that renders the same signal more accurately can agree to; it justifies the tolerance of
tests/test_round2_gpu.py::test_material_sync_train_inner_step.  Run in the build container:

    python -m oracle.sync_grad_sensitivity        # prints the relative differences, writes tests/golden/step_material_sync_f64.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                  # noqa: E402
from oracle import make_goldens_r2 as g2              # noqa: E402


def main():
    assert rh.reference_available()
    R = rh.ref_modules()
    ms = g2.ref_mss()
    dm, osc = R.diff_model, R.oscillator
    Material = R.material_model.Material
    g = np.load(os.path.join(ROOT, "tests", "golden", "step_material_sync.npz"))
    sample_rate, frame_num, force_frame_num, eigen_num = (int(x) for x in g["meta"])
    init_coeff = [float(x) for x in g["init_coeff"]]
    gt_forces = torch.zeros((1, force_frame_num))
    gt_forces[0, 0] = 1
    out = {}
    for dt in (torch.float32, torch.float64, 'again'):
        key = dt
        if dt == 'again':
            dt = torch.float32
        model = dm.build_model(g2.bowl_mesh_dir(), mode_num=eigen_num, order=1, mat=init_coeff, task="material")
        mm = model.material_model
        with torch.no_grad():
            mm.youngs.probablity.copy_(torch.tensor(g["youngs_logits0"]))
            mm.poisson.probablity.copy_(torch.tensor(g["poisson_logits0"]))
        oscillator = osc.TraditionalDampedOscillator(gt_forces.to(dt), 1, eigen_num, frame_num, sample_rate, Material(init_coeff))
        late = ms.MSSLoss([1024, 512, 256, 128, 64], sample_rate, type="l1_loss").to(dt)
        model.eigen_decomposition()
        undamped = model.get_undamped_freqs().to(dt)
        undamped.retain_grad()
        predict = oscillator(undamped)
        loss = late(predict, torch.tensor(g["gt_audios"]).to(dt), oscillator.damped_freq, 1)
        loss.backward()
        out[key] = dict(lam=model.eigenvalues.numpy().copy() if hasattr(model, 'eigenvalues') else None, loss=float(loss), gE=mm.youngs.probablity.grad.numpy().copy(), gnu=mm.poisson.probablity.grad.numpy().copy(),
                       gf=undamped.grad.numpy().copy(), predict=predict.detach().numpy().copy())
    a, b = out[torch.float32], out[torch.float64]
    rel = lambda x, y: float(np.linalg.norm(np.asarray(x, np.float64) - np.asarray(y, np.float64)) / np.linalg.norm(np.asarray(y, np.float64)))
    print("fp32 run vs committed golden: loss", a["loss"], float(g["loss"]), "gE", rel(a["gE"], g["grad_youngs_logits"]))
    c = out["again"]
    print("reference fp32, second run in the same process vs first: dL/d(youngs logits)", rel(c["gE"], a["gE"]), " dL/d(poisson logits)",
          rel(c["gnu"], a["gnu"]), " dL/dfreq", rel(c["gf"], a["gf"]))
    print("second run vs committed golden: gE", rel(c["gE"], g["grad_youngs_logits"]), "gnu", rel(c["gnu"], g["grad_poisson_logits"]))
    print("first run vs committed golden: gnu", rel(a["gnu"], g["grad_poisson_logits"]))
    if a["lam"] is not None:
        lam = np.sort(np.asarray(a["lam"]).ravel())
        print("  relative gaps of consecutive eigenvalues:", np.array2string((lam[1:] - lam[:-1]) / lam[1:], precision=2))
    print("reference fp32 vs reference fp64:")
    print("  audio rel-L2", rel(a["predict"], b["predict"]))
    print("  loss", a["loss"], b["loss"], abs(a["loss"] - b["loss"]) / b["loss"])
    print("  dL/dfreq rel-L2", rel(a["gf"], b["gf"]))
    print("  dL/d(youngs logits)", rel(a["gE"], b["gE"]), " dL/d(poisson logits)", rel(a["gnu"], b["gnu"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "step_material_sync_f64.npz"), loss=np.array(b["loss"]),
                        grad_youngs_logits=b["gE"], grad_poisson_logits=b["gnu"], grad_freq=b["gf"], predict=b["predict"].astype(np.float32),
                        grad_freq_f32=a["gf"])


if __name__ == "__main__":
    main()
