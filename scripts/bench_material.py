"""BASELINE.json configs[1]: material inference step on a quadratic mesh, 1 x B200.
One epoch step of experiments/material_sync_train.py mode 3: eigen-decomposition once (the experiment does it once
per object), then per step get_undamped_freqs() forward + TraditionalDampedOscillator (1 audio, 16 modes, 8 000
samples, sr 32 000, 150-tap impulse force) forward, a spectral-free L2 loss on the audio, backward to the 16 + 16
Young / Poisson logits.  Meshes: the bowl and grid16 fixtures of tests/golden/meshes.npz, order 2.
Prints one JSON line per mesh (CUDA-event times)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffsound_b200.diffelastic.diff_model import DiffSoundObj, TrainableLinear
from diffsound_b200.diffelastic.material_model import Material, MatSet
from diffsound_b200.ddsp.oscillator import TraditionalDampedOscillator

dev = torch.device("cuda:0")
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes.npz"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for name in ("bowl", "grid16"):
    v, t = d[f"{name}_verts"], d[f"{name}_tets"].astype(np.int64)
    k, T, sr = 16, 8000, 32000
    force = torch.zeros(1, 150, device=dev)
    force[0, 0] = 1
    for rep in range(2):
        e0 = ev()
        obj = DiffSoundObj(torch.tensor(v, device=dev), torch.tensor(t, device=dev), mode_num=k, order=2,
                           mat=MatSet.Ceramic, mat_model=TrainableLinear, task="material")
        obj.eigen_decomposition()
        e1 = ev()
    osc = TraditionalDampedOscillator(force, 1, k, T, sr, Material(MatSet.Ceramic))
    params = list(obj.material_model.parameters())
    with torch.no_grad():
        target = osc(obj.get_undamped_freqs().detach() * 1.03)

    def step():
        for p in params:
            p.grad = None
        f = obj.get_undamped_freqs()
        y = osc(f)
        loss = ((y - target) ** 2).mean()
        loss.backward()
        return loss

    for _ in range(3):
        step()
    s0 = ev()
    for _ in range(reps):
        loss = step()
    s1 = ev()
    torch.cuda.synchronize()
    pat = obj.deform.pattern
    print(json.dumps({"what": "material inference step (config 2)", "mesh": name, "order": 2, "tets": int(t.shape[0]),
                      "n": pat.n, "nnz": pat.nnz, "modes": k, "setup_plus_eigen_ms": e0.elapsed_time(e1),
                      "step_ms": s0.elapsed_time(s1) / reps, "steps_per_s": 1e3 * reps / s0.elapsed_time(s1),
                      "lobpcg": obj.eig_stats, "loss": float(loss),
                      "grad_norm": float(sum(float((p.grad ** 2).sum()) for p in params) ** 0.5)}), flush=True)
