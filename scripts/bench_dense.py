"""FP64 dense kernels of a LOBPCG step at the bench size (n = 823 875, m = 48): Gram strips and the lean Ritz update.

    DS_STRIP_CFG=1651 python scripts/bench_dense.py      # ring geometry of k_gram_strip: rows * 100 + stages * 10 + producer
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsound_b200 import native  # noqa: E402


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
    n, m = 823875, 48
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    S = torch.randn(n, 3 * m, dtype=torch.float64, device=dev, generator=g)
    KS = torch.randn(n, 3 * m, dtype=torch.float64, device=dev, generator=g)
    MS = torch.randn(n, 3 * m, dtype=torch.float64, device=dev, generator=g)
    out = {"what": "LOBPCG dense kernels", "n": n, "m": m, "strip_cfg": os.environ.get("DS_STRIP_CFG", "880")}
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fp64_peak.json")))
    for wa in (48, 32, 16):
        KW, MW = KS[:, m:m + wa], MS[:, m:m + wa]
        GsK, GsM = native.gram_strip(KW, MW, S)
        ref = KW[:65536].T @ S[:65536]
        chk_K, _ = native.gram_strip(KS[:65536, m:m + wa], MS[:65536, m:m + wa], S[:65536])
        err = float((chk_K - ref).abs().max() / ref.abs().max())
        ms = timed(lambda: native.gram_strip(KW, MW, S))
        fl = 2.0 * n * (2 * wa) * (3 * m)
        out[f"gram_strip_wa{wa}"] = {"ms": ms, "TFLOPs": fl / ms / 1e9, "frac_dmma_peak": fl / ms / 1e9 / peak["dmma_m8n8k4_tflops"],
                                     "GBps": n * (2 * wa + 3 * m) * 8.0 / ms / 1e6, "rel_err_64k_rows": err}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
