"""Modal synthesis at BASELINE.json configs[4]: 256 modes x 44.1 kHz x 2 s x batch 1024
damped sinusoids, forward and backward (gradients w.r.t. amplitudes, damping, frequencies; upstream gy = y).
CUDA-event times, achieved FP32 rate against the FFMA peak of the SMs, and a parity check of a batch slice against
the fp64 closed form of the reference formula (oscillator.py:297-304).  Prints one JSON line.

Under torchrun (WORLD_SIZE > 1) the batch is sharded over the ranks (diffsound_b200/parallel/synth.py): every rank renders
B / world rows, the backward pass all-reduces the 2 x k shared gradients over NCCL, the time is the max over ranks and
the rates are whole-job aggregates.

usage: python scripts/bench_synth.py [B] [k] [T] [reps]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_synth.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffsound_b200 import native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = int(sys.argv[3]) if len(sys.argv) > 3 else 88200
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
sr = 44100.0
import torch.distributed as dist
from diffsound_b200.parallel.synth import allreduce_shared_grads, batch_slice
rank, world, local = (int(os.environ.get(k_, d_)) for k_, d_ in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
B_total = B
g = torch.Generator().manual_seed(0)
f = torch.sort(torch.rand(k, generator=g) * (18000 - 100) + 100).values
alpha = torch.exp(torch.rand(k, generator=g) * np.log(100.0) + np.log(0.6))
beta = torch.exp(torch.rand(k, generator=g) * np.log(100.0) + np.log(1e-8))
d = 0.5 * (alpha + beta * (2 * np.pi * f) ** 2)
fd = torch.sqrt(torch.clamp((2 * np.pi * f) ** 2 - d ** 2, min=0.0)) / (2 * np.pi)
amp = 2 * torch.sigmoid(torch.rand(B, k, generator=g) * 0.04) ** 2.3 + 1e-6      # ddsp/utils.py:6-9 modifed_sigmoid
amp = amp[batch_slice(B_total, rank, world)]
B = amp.shape[0]
amp_d, d_d, f_d = amp.float().to(dev), d.float().to(dev), fd.float().to(dev)


def timed(fn):
    for _ in range(2):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


ms_f, y = timed(lambda: native.modal_synth_fwd(amp_d, d_d, f_d, T, sr))
def bwd():
    ga, gd, gf = native.modal_synth_bwd(amp_d, d_d, f_d, y, sr)
    allreduce_shared_grads(gd, gf)              # the path's only collective: 2 x k floats
    return ga, gd, gf


ms_b, grads = timed(bwd)
# parity on a batch slice, fp64 closed form
nb = 4
tau = (np.arange(T, dtype=np.float64) + 1.0) / sr
dd, ff = d_d.cpu().double().numpy(), f_d.cpu().double().numpy()
basis = np.exp(-dd[:, None] * tau[None, :]) * np.sin(2 * np.pi * ff[:, None] * tau[None, :])
ref = amp_d[:nb].cpu().double().numpy() @ basis
err = float(np.linalg.norm(y[:nb].cpu().double().numpy() - ref) / np.linalg.norm(ref))
gamp_ref = y[:nb].cpu().double().numpy() @ basis.T
gerr = float(np.linalg.norm(grads[0][:nb].cpu().double().numpy() - gamp_ref) / np.linalg.norm(gamp_ref))
sms = torch.cuda.get_device_properties(0).multi_processor_count
peak_tf = sms * 128 * 2 * 1.965e9 / 1e12          # FFMA lanes x 2 flop x boost clock
ms_samples = B_total * k * T
peak_tf *= world
line = {"what": "modal synthesis (config 5)", "n_gpus": world, "scaling": "strong (batch sharded)", "B": B_total, "B_per_gpu": B,
        "modes": k, "T": T, "sr": sr,
        "fwd_ms": ms_f, "bwd_ms": ms_b, "fwd_G_mode_samples_per_s": ms_samples / ms_f / 1e6,
        "fwd_TFLOPs_contraction": 2 * ms_samples / ms_f / 1e9, "bwd_TFLOPs_contraction": 4 * ms_samples / ms_b / 1e9,
        "fp32_peak_TFLOPs": peak_tf, "fwd_frac_of_fp32_peak": 2 * ms_samples / ms_f / 1e9 / peak_tf,
        "bwd_frac_of_fp32_peak": 4 * ms_samples / ms_b / 1e9 / peak_tf,
        "output_GB_per_s": B_total * T * 4 / ms_f / 1e6, "audio_rel_l2_vs_fp64": err, "gamp_rel_l2_vs_fp64": gerr,
        "note": "flops counted for the batch x mode x time contraction only (2 per mode-sample forward; backward = 2 "
                "contractions: gamp and z = A^T gy); the basis recurrence adds 8 flop per (mode, sample) per 64-row batch tile"}
if rank == 0:
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
