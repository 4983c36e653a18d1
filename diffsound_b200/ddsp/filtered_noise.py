"""DDSP filtered noise.  API mirror of src/ddsp/filtered_noise.py:7-67 (`FilteredNoise`): same
constructor arguments, the same `coefficient_bank` parameter (shape, init) and output shape.

The reference builds the time-varying filter bank and convolves with five FFTs and a
`conv_transpose1d` overlap-add; here both passes are one kernel each in the time domain
(`ds_filtered_noise_fwd` / `_bwd`, csrc/noise.cu).  The white-noise frames are drawn exactly as the
reference draws them -- `torch.rand` on the host generator, then moved to the device -- so a seeded run
reproduces the reference's noise."""
import torch
import torch.nn as nn

from .. import native


class _FilteredNoiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeff, noise, sample_num, gain):
        c = coeff.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(c, noise)
        ctx.gain = gain
        ctx.in_dtype = coeff.dtype
        return native.filtered_noise_fwd(c, noise, sample_num, gain)

    @staticmethod
    def backward(ctx, gy):
        c, noise = ctx.saved_tensors
        return native.filtered_noise_bwd(c, noise, gy, ctx.gain).to(ctx.in_dtype), None, None, None


class FilteredNoise(nn.Module):
    def __init__(self, noise_num, sample_num, filter_coeff_length=65, frame_length=64, attenuate_gain=1.0, device="cuda"):
        super().__init__()
        self.frame_length = frame_length
        self.filter_coeff_length = filter_coeff_length
        self.noise_num = noise_num
        self.sample_num = sample_num
        self.device = device
        self.attenuate_gain = attenuate_gain
        # drawn on the host generator like the reference's, then placed on `device` when there is one (the
        # reference's scripts move the whole oscillator with .cuda() right after construction)
        bank = torch.zeros(noise_num, sample_num // frame_length + 1, filter_coeff_length).uniform_(-1, 1)
        if torch.cuda.is_available() and str(device).startswith("cuda"):
            bank = bank.to(torch.device(device) if str(device) != "cuda" else torch.device("cuda", torch.cuda.current_device()))
        self.coefficient_bank = nn.Parameter(bank)

    def forward(self):
        x = self.coefficient_bank
        if not x.is_cuda:
            raise RuntimeError("diffsound_b200: FilteredNoise must live on a CUDA device (call .cuda(); there is no CPU path)")
        batch_num, frame_num, _ = x.shape
        # filtered_noise.py:47-48: host generator, then the device
        noise = torch.rand(batch_num, frame_num, self.frame_length, dtype=torch.float32).to(x.device) * 2 - 1
        return _FilteredNoiseFn.apply(x, noise.contiguous(), self.sample_num, float(self.attenuate_gain))
