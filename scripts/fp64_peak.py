"""Measured FP64 peaks of this device (register-resident loops, csrc/rr.cu k_fp64_peak): the denominators of the
FP64 rooflines (DMMA m8n8k4 for the Gram / update kernels, DFMA for the SpMM accumulation).  Writes one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsound_b200 import native

res = {"device": torch.cuda.get_device_name(0)}
for mode, name in ((0, "dfma"), (1, "dmma_m8n8k4")):
    best = 0.0
    for cps in (1, 2, 4):
        for _ in range(3):
            best = max(best, native.fp64_peak(mode, 8192, cps))
    res[name + "_tflops"] = best
res["how"] = ("512-thread CTAs, 16 independent accumulator chains per thread (DFMA) / 8 independent m8n8k4 accumulators per warp "
              "(DMMA), 8192 iterations, CUDA events, best of 9 launches over 1/2/4 CTAs per SM")
print(json.dumps(res))
