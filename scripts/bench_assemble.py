"""Assembly of K and M on the bench mesh (32^3 x 6 quadratic Kuhn cube): the two row kernels, CUDA-event times.

    python scripts/bench_assemble.py [cells_per_side]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffsound_b200 import native  # noqa: E402
from diffsound_b200.diffelastic import mass_matrix as mmx  # noqa: E402
from diffsound_b200.diffelastic.diff_model import DiffSoundObj  # noqa: E402


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    v, t = bench.kuhn_cube(N)
    obj = DiffSoundObj(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev), mode_num=32, order=2, mat=bench.STEEL)
    d = obj.deform
    verts, tets = d.verts_f32(), d.tets_i32
    pat = d.pattern
    mu, lam = 7.7e10, 1.1e11
    ctab, mtab = mmx.stiffness_contraction_table(2).to(dev), mmx.mass_density_table(2, 7850.0).to(dev)
    Kval = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
    Mblk = torch.empty(pat.nnzb, dtype=torch.float64, device=dev)
    geom = torch.empty(tets.shape[0] * 14, dtype=torch.float64, device=dev)
    out = {"what": "assembly K + M (geometry + row kernel)", "tets": int(tets.shape[0]), "nodes": pat.n_nodes, "nnzb": pat.nnzb,
           "max_deg": pat.max_deg}
    survey_bytes = 2.0 * 9.0 * pat.nnzb * 8.0 + tets.numel() * 4.0 + pat.n_nodes * 12.0 + tets.shape[0] * 100 * 4.0
    written = pat.nnzb * 80.0
    for k in ("tets", "rows"):
        ms = timed(lambda: native.assemble_km(verts, tets, 2, pat, mu, lam, ctab, mtab, Kval=Kval, Mblk=Mblk, geom=geom, kernel=k))
        out[k] = {"ms": ms, "GBps_survey_bytes": survey_bytes / ms / 1e6, "GBps_written": written / ms / 1e6}
    out["pattern_ms_with_slot"] = timed(lambda: native.Pattern(tets, pat.n_nodes, want_slot=True), reps=5)
    out["pattern_ms_without_slot"] = timed(lambda: native.Pattern(tets, pat.n_nodes, want_slot=False), reps=5)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
